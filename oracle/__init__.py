"""CPU oracle for the FEM elasto-plasticity hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``fem_elastoplasticity_b200`` imports this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may use it, and only as the checker or
the CPU comparator - never as the thing shipped.

Parity status: PINNED.  ``oracle.fem_oracle`` is checked bit-for-bit against the
reference's own functions imported from ``/root/reference`` (``tests/test_oracle_vs_reference.py``,
container only), against fixtures generated from the reference by
``oracle/make_golden.py`` and committed under ``tests/golden/`` (these travel to
the GPU box), and against the reference's tsx-tunnel golden CSVs
(``k_tangent_qq.csv``, ``fq.csv``, ``f0q.csv``; condensed into
``tests/golden/tsx_csv_golden.npz``).
"""
