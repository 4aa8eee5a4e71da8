"""CPU oracle for the GarmentNets dense-inference hot path.

TEST INFRASTRUCTURE ONLY.  This package restates, on the CPU, the algorithm of every reference function on the hot
path so that the CUDA implementation can be checked against it.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it; nothing
under ``garmentnets_b200/`` does.

Pinning status:
  * ATen-backed stages (MLP, 3D-UNet, grid_sample decoder) and ``VirtualGrid`` / ``ArraySlicer`` are PINNED: the
    oracle is compared with the reference's own importable modules (``components/unet3d.py``, ``components/mlp.py``,
    ``components/gridding.py``) by ``oracle/make_golden.py``, whose outputs are committed under ``tests/golden/``.
  * ``gaussian_gradient_magnitude`` is PINNED against scipy.ndimage (present in the image).
  * fps / radius / knn_interpolate / PointConv / scatter (torch_cluster 1.5.9, torch_geometric 1.7.2, torch_scatter
    2.0.8) and Lewiner marching cubes (scikit-image 0.18.2) are third-party binaries that are absent from
    /root/reference and from this image: PARITY UNPINNED.  The oracle restates their published semantics; ball
    query / kNN sets are cross-checked with scipy.spatial.cKDTree, marching cubes by topological invariants.
"""
