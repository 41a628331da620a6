"""CPU oracle: a NumPy restatement of the Monte-Carlo hot path of rnissel/Channel-Estimation.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this.  Only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may use it, and only as the checker / CPU baseline.

PARITY UNPINNED: the reference ships no tests, golden vectors, fixtures or seeds, and
neither MATLAB nor GNU Octave exists in this image, so the reference itself cannot be run.
This restatement follows the reference line by line (every function cites the
``file:line`` it follows) and is pinned only by the identities the reference states in its
comments, by its closed-form theory curve and by hand-derived constants (see
``tests/test_oracle_*.py``).
"""
