"""CPU oracle for the gym-auv per-step hot path.  TEST INFRASTRUCTURE ONLY.

This package is a plain FP64 NumPy/Python restatement of the reference's
``BaseEnvironment.step()`` path (``/root/reference/gym_auv/environment.py:292-366``
and everything it calls).  It exists to *check* the CUDA path; it is never the
thing that is shipped or measured as the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  Nothing under ``gym_auv_b200/`` imports it.

Parity status
-------------
* dynamics (``odesolver45`` + ``_state_dot`` + model matrices), ``princip`` and the
  sector map are PINNED: ``tests/golden/make_reference_goldens.py`` runs the
  reference's own NumPy-only files (``utils/constants.py``, ``utils/geomutils.py``,
  ``objects/vessel/odesolver.py``, ``utils/sector_partitioning.py``) and commits
  their outputs as fixtures; ``tests/test_oracle_pinned.py`` checks this
  restatement against them bit-for-bit (dynamics to 1e-15).
* the reference's own classes run behind import stubs
  (``tests/golden/make_reference_goldens_stubbed.py`` -> ``reference_stubbed.npz``) PIN the
  culling-window integers incl. Python's negative-index wrap (``sensor.py:41-97``), feasibility
  pooling, ``Path`` / ``RandomCurveThroughOrigin``, ``Vessel.step`` + ``Vessel.navigate`` rollouts,
  ``VesselObstacle`` tracks and both rewarders (``tests/test_reference_goldens_stubbed.py``).
* ``tests/golden/make_reference_goldens_hybrid.py`` runs the reference's LiDAR pipeline classes
  unmodified on ``geos_lite`` primitives and pins the glue (nearby list, per-ray obstacle lists,
  min over intersection pieces, closeness, collision, obs, reward) -- not the primitives.
* what the reference computes INSIDE Shapely 1.7.0 / GEOS (ray/boundary
  intersection, ``Point.distance``, ``LineString.project``, ``buffer().simplify()``,
  ``minimum_rotated_rectangle``, ``affinity.rotate``) is restated from GEOS' published
  algorithms in ``oracle/geos_lite.py`` -- Shapely/GEOS is not installable here
  (no wheel, no libgeos, no network) so that part is **PARITY UNPINNED**: it is
  anchored only on the reference's own qualitative test
  (``tests/test_hierarchical_collision_detector.py:38-48``) and on independent
  geometric cross-checks (analytic circle vs polygonised circle, brute-force vs
  culled casting).
"""
