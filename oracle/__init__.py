"""CPU oracle for the HIAST post-logit self-training hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``hiast_b200/`` may import this package.
The only callers allowed are ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` -- and there only as
the checker or the timed CPU baseline, never as the shipped path.

What it is: a numpy / torch-CPU restatement of the reference algorithms
(bupt-ai-cz/HIAST, mounted read-only at /root/reference while this repo was
built).  Each function cites the reference ``file:line`` it follows (paths are
relative to ``/root/reference/code``).

Parity pinning: the reference ships no tests, golden vectors or fixtures
(SURVEY.md section 4), so the oracle is pinned against outputs of the UNMODIFIED
reference code executed in the build container: ``tests/golden/make_golden.py``
imports the reference through a stub shim (apex / tensorboardX / albumentations
are absent), runs it on seeded inputs and commits the results as
``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` checks every oracle
function against those fixtures bit-for-bit (losses: to 1e-6).

Third-party arithmetic the reference leans on and that is therefore part of the
contract (installed versions are the ground truth):
numpy 2.3.5 ``np.quantile(method='linear')``, ``np.mean``, ``astype(float16)``,
scalar ``**`` (glibc pow); torch 2.11 ``softmax`` / ``max`` / ``log_softmax`` /
``CrossEntropyLoss`` / ``histc``.
"""

from . import ias, losses, metrics, copy_paste, ema  # noqa: F401
