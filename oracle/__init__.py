"""CPU oracle for the DeepFwFM forward hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
package ``xsdeepfwfm_deprecated_b200``; only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it, and
there only as the checker (or as the timed CPU baseline), never as the thing shipped.

Parity status: PINNED.  The reference ships no golden vectors or tests of its own
(SURVEY.md section 4), so the pins are outputs of the reference module itself,
imported unmodified from ``/root/reference`` in the build container by
``tests/golden/make_golden.py`` and committed as fixtures under ``tests/golden/``.
``tests/test_oracle_golden.py`` checks both restatements below against every one
of those fixtures.

Modules
-------
closed_form   fp64 numpy closed form of ``DeepFMs.forward`` (model/DeepFMs.py:285-469)
torch_port    fp32 torch-CPU restatement that follows the reference's op sequence
              (same materialised F x F x B x K outer products) -- the timed CPU baseline
prune         the reference's magnitude-threshold bisection and the one-shot recipe
synth         deterministic synthetic weights / inputs / cardinalities
"""
