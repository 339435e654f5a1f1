"""CPU oracle for the scanpath sampling + scoring hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``scanpaths_b200/`` may import this
package: it is the checker for the CUDA path, never the thing shipped or
measured.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it.

Every function is an independent restatement of the reference algorithm
(chenxy99/Scanpaths) and cites the reference file:line it follows.  Parity of
this oracle is PINNED against the reference itself: ``tests/golden/make_goldens.py``
imports the unmodified reference from ``/root/reference`` (with import stubs for
matplotlib / multimatch_gaze / mmcv) and records its outputs into
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks the oracle against
those fixtures, and ``tests/test_oracle_vs_reference.py`` re-checks it live
whenever ``/root/reference`` is mounted.

MultiMatch (external ``multimatch-gaze==0.1.2``, not vendored by the reference,
not installed here) is out of scope: parity unpinned for those 5 numbers only.
"""
