"""ml_qem_b200 -- B200-native exact expectation-value engine for ML-QEM (qiskit-community/ml-qem).

Drop-in for the reference's Aer-backed Estimator on its data-generation hot path
(blackwater/data/utils.py:418-444): ``B200Estimator.run(circuits, observables, parameter_values)
-> job -> EstimatorResult(values, metadata)``.  CUDA kernels (sm_100a) behind a C ABI
(include/bwq.h); no CPU fallback.
"""
from .circuit import Circuit, Parameter, parse_qasm  # noqa: F401
from .observable import PauliObservable  # noqa: F401
from .backends import BackendProps, fake_belem, fake_lima, fake_montreal, synthetic_chain  # noqa: F401
from .noise import NoiseModel, from_backend as noise_from_backend  # noqa: F401

__version__ = "0.1.0"


def __getattr__(name):  # lazy: the estimator pulls in ctypes + the CUDA library
    if name in ("B200Estimator", "EstimatorResult", "B200Job"):
        from . import estimator

        return getattr(estimator, name)
    if name in ("Engine", "EngineError", "FlatBatch", "encode_batch", "Variants"):
        from . import engine

        return getattr(engine, name)
    raise AttributeError(name)
