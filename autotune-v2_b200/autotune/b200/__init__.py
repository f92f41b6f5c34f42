"""B200 (sm_100a) back end of the autotune drop-in: C-ABI binding (`_native`), torch custom ops (`ops`, registered as
`torch.ops.at2.*` on import) and the batched entry points (`engine`, `epdf`)."""
from . import ops  # noqa: F401  (registers torch.ops.at2.*; raises if libat2_b200.so is missing - no CPU fallback)
