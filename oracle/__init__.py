"""CPU oracle for the open-pcc-metric hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it, and only as the checker.  The product package
(``open_pcc_metric_b200``) never imports this directory and fails loudly when its
CUDA library is missing.

Parity status: **parity unpinned at the Open3D boundary.**  The reference
(/root/reference/open_pcc_metric) delegates all geometry to Open3D 0.18.0
(pyproject.toml:12, poetry.lock:1131-1132), which is neither vendored in the
reference tree nor installable in this environment, and the reference's own
tests (tests/unit/test_metric.py) pin no number at that boundary.  What *is*
pinned:

* the pure-Python part of the reference (cloud_pair.py, metric.py, calculator.py,
  options.py) is executed UNMODIFIED, in the build container, over the Open3D
  stand-in in ``o3d_standin.py`` (``make_golden.py``); its outputs are committed
  under ``tests/golden/`` and the numpy port in ``reference_port.py`` is checked
  against them bit for bit;
* the three facts the reference tests do assert (test_metric.py:44-47, :65-70)
  and the deterministic fixture it defines (test_metric.py:13-26, "KA-1").

Modules
-------
o3d_standin     Open3D 0.18.0 stand-in: PointCloud, KDTreeFlann, Vector3dVector,
                estimate_normals, compute_nearest_neighbor_distance,
                get_minimal_oriented_bounding_box (published algorithms restated).
reference_port  Vectorised numpy restatement of cloud_pair.py + metric.py.
cnn             ctypes loader for the plain-C brute-force NN/kNN (nn_brute.c).
make_golden     Script that runs the unmodified reference here and writes
                tests/golden/*.
"""
