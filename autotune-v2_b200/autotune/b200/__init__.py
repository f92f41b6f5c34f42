"""B200 (sm_100a) back end of the autotune drop-in: C-ABI binding, torch custom ops and the batched entry points."""
