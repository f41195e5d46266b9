"""Import shim: `import flinkms_b200` loads the package directory `flink-ms_b200/` (hyphenated name)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("flink-ms_b200")
sys.modules[__name__] = _pkg

if __name__ == "__main__":                      # python flinkms_b200.py als|svm ...  (or python -m flinkms_b200 als|svm ...)
    sys.exit(importlib.import_module("flink-ms_b200.jobs").main())
