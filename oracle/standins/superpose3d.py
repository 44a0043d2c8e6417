"""Stand-in for ``superpose3d`` (absent): utils/modeler.py:26 imports it at module level;
nothing on the N1 path (Solver.clustering) calls it."""


def Superpose3D(*a, **k):
    raise NotImplementedError("stand-in: not on the tested path")
