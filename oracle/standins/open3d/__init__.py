"""Stand-in for ``open3d`` (absent from this image), used ONLY by oracle/ref_harness.py to
run the unmodified ``Solver.clustering`` (utils/modeler.py:762-899).

The reference needs three names: ``o3d.geometry.PointCloud`` with a ``points`` attribute,
``o3d.utility.Vector3dVector`` and ``PointCloud.cluster_dbscan(eps, min_points)``
(utils/modeler.py:768-770).  DBSCAN is restated from its published definition (Ester et al.
1996) with Open3D's conventions: the eps-neighbourhood of a point includes the point itself
and is closed (distance <= eps), a point is a core point when that neighbourhood holds at
least ``min_points`` points, clusters are numbered in the order their first core point
appears in the input, a border point joins the first cluster that reaches it, noise is -1.
scikit-learn's DBSCAN (installed) follows the same conventions and is what runs here.
Open3D itself is not available, so this step is "parity unpinned" against the real package.
"""
import types

import numpy as np


def _vector3d(a):
    return np.asarray(a, dtype=np.float64).reshape(-1, 3)


class _PointCloud:
    def __init__(self):
        self.points = np.zeros((0, 3))

    def cluster_dbscan(self, eps, min_points, print_progress=False):
        from sklearn.cluster import DBSCAN
        pts = np.asarray(self.points, dtype=np.float64)
        if len(pts) == 0:
            return []
        return DBSCAN(eps=float(eps), min_samples=int(min_points), algorithm='kd_tree').fit(pts).labels_.tolist()


geometry = types.SimpleNamespace(PointCloud=_PointCloud)
utility = types.SimpleNamespace(Vector3dVector=_vector3d)
