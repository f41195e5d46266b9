"""flink-ms_b200 -- B200-native ALS / CoCoA-SVM hot paths behind the FlinkML operator surface that
mmziyad/flink-ms programs against.  The directory name carries a hyphen (it mirrors the upstream repo
name), so import it as `import flinkms_b200` (shim at the repo root) or
`importlib.import_module("flink-ms_b200")`.
"""
from . import _native
from ._native import Context, FlinkMSError
from .operators import ALS, SVM, DenseVector, Factors, OnlineSGD, ParameterMap, readCsvFile, readLibSVM

__all__ = ["ALS", "SVM", "DenseVector", "Factors", "OnlineSGD", "ParameterMap", "readCsvFile", "readLibSVM", "Context", "FlinkMSError", "_native"]
