"""qcqp_b200 -- B200-native engine behind the Suggest-and-Improve hot path of cvxgrp/qcqp.

Same surface as the reference package (qcqp/__init__.py:27-29): QCQP and the method-name constants.
Importing the package loads libqcqp_b200.so and fails loudly when it is missing; there is no CPU fallback."""
from . import _lib

_lib.load()

from .qcqp import QCQP                                           # noqa: E402
from .settings import RANDOM, SPECTRAL, SDR                      # noqa: E402
from .settings import COORD_DESCENT, ADMM, DCCP, IPOPT           # noqa: E402
from .forms import QuadraticFunction, QCQPForm                   # noqa: E402

__all__ = ["QCQP", "RANDOM", "SPECTRAL", "SDR", "COORD_DESCENT", "ADMM", "DCCP", "IPOPT", "QuadraticFunction", "QCQPForm"]
