"""CPU oracle for the MCOS Monte-Carlo loop -- TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU with numpy / scipy / scikit-learn, the algorithms of the reference's hot path
(`mcos/mcos.py:13-64` and the plugin classes it calls).  It exists so that the CUDA engine in `mcos_b200/` can be
checked against the reference's numbers on identical injected inputs.

Rules (enforced by review):
  * only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py` may import
    anything from here;
  * nothing under `mcos_b200/` imports it -- the product path has no CPU fallback.

Parity pinning: every function here is checked by `tests/test_oracle_golden.py` against the golden vectors that the
reference's own tests hold (`tests/test_optimizer.py`, `tests/test_covariance_transformer.py`,
`tests/test_error_estimator.py`, `tests/test_utils.py`) and against outputs of the reference modules themselves run in
the build container (`tests/golden/make_golden.py` -> `tests/golden/*.npz`).  The one piece that is NOT pinned is
NCO's KMeans clustering step (sklearn 0.22.1's k-means++ stream cannot be reproduced with the installed sklearn 1.9):
"parity unpinned" for `nco_cluster`, pinned for the NCO allocation algebra given a partition.
"""
from .mcos_port import *  # noqa: F401,F403
