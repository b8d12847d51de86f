"""B200-native hot path of AdditiveCausalExpansion (R package `ace` 0.4.1): the empirical-Bayes GP step
(additive kernel build -> Cholesky/inverse -> log-evidence + hyper-parameter gradients -> Nadam) and the
posterior, as hand-written sm_100a CUDA behind the reference's own native interface.

* `api`     -- the reference's exported native routines, same names (R/RcppExports.R)
* `fit`     -- device-resident fit handle (`AceFit`): the body of Kernel$para_update / predict
* `kernel`  -- mirrors of the R6 kernel / optimiser classes and of ace.train / predict.ace
* `basis`   -- treatment bases (inputs of the path)
* `synth`   -- synthetic inputs of the BASELINE.json shapes

No CPU fallback: the compute entry points need libace_b200.so and a B200.
"""
from . import _lib  # noqa: F401

__all__ = ["api", "fit", "kernel", "basis", "synth"]
