class Axes3D:
    name = '3d'
