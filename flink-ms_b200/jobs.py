"""The two training jobs of the reference as drop-in drivers over the GPU operators.

`als_impl(argv)` mirrors `flink-als/src/main/scala/de/tub/it4bi/ALSImpl.scala:15-75` and `svm_impl(argv)` mirrors
`flink-svm/src/main/scala/de/tub/it4bi/SVMImpl.scala:13-56`: same `--key value` flags (Flink's ParameterTool), same
defaults, same text rows on stdout or in the output files.  The inputs are parsed on the GPU (row N4), the models are
trained by the CUDA kernels behind `operators.ALS` / `operators.SVM`, the rows are formatted like Java does
(`formats.py`).  Run as  `python -m flinkms_b200 als --input ratings.csv --numFactors 64 ...`  (the import shim at the repo root dispatches here).

Differences a user should know: `--blocks` / `--temporaryPath` of the ALS job are accepted and have no effect (no Flink
shuffles here); an output path is written as ONE text file (Flink writes a directory of part files at parallelism > 1);
the SVM block of a row is `row_index * parallelism // n % blocks` with `--parallelism` (default 1) standing in for the
Flink source parallelism that decides it in the reference.
"""
from __future__ import annotations

import sys

import numpy as np

from . import _native as N
from .formats import als_factor_row, svm_model_rows
from .operators import ALS, SVM, ParameterMap, readCsvFile, readLibSVM


class ParameterTool:
    """org.apache.flink.api.java.utils.ParameterTool.fromArgs: `--key value` / `-key value` pairs, a key without a value
    maps to the marker Flink calls NO_VALUE_KEY ("__NO_VALUE_KEY")."""

    NO_VALUE_KEY = "__NO_VALUE_KEY"

    def __init__(self, data: dict):
        self.data = dict(data)

    @staticmethod
    def fromArgs(args):
        data, k = {}, 0
        args = list(args)
        while k < len(args):
            a = args[k]
            if a.startswith("--"):
                key = a[2:]
            elif a.startswith("-"):
                key = a[1:]
            else:
                raise ValueError(f"Error parsing arguments '{args}' on '{a}'. Please prefix keys with -- or -.")
            if not key:
                raise ValueError("The input " + str(args) + " contains an empty argument")
            k += 1
            if k >= len(args):
                data[key] = ParameterTool.NO_VALUE_KEY
            else:
                nxt = args[k]
                is_number = False
                try:
                    float(nxt); is_number = True
                except ValueError:
                    pass
                if nxt.startswith("-") and not is_number:
                    data[key] = ParameterTool.NO_VALUE_KEY
                else:
                    data[key] = nxt
                    k += 1
        return ParameterTool(data)

    def has(self, key):
        return key in self.data

    def get(self, key, default=None):
        return self.data.get(key, default)

    def getRequired(self, key):
        if key not in self.data:
            raise RuntimeError(f"No data for required key '{key}'")
        return self.data[key]

    def getInt(self, key, default=None):
        return int(self.data[key]) if key in self.data else default

    def getLong(self, key, default=None):
        return self.getInt(key, default)

    def getDouble(self, key, default=None):
        return float(self.data[key]) if key in self.data else default

    def getBoolean(self, key, default=False):
        return self.data[key].lower() == "true" if key in self.data else default     # Boolean.valueOf


def _write_text(path, rows):
    with open(path, "w") as f:                                                      # WriteMode.OVERWRITE
        for r in rows:
            f.write(r + "\n")


def als_impl(argv, ctx: N.Context | None = None, out=sys.stdout):
    """ALSImpl.main(args)."""
    params = ParameterTool.fromArgs(argv)
    delim_name = params.get("fieldDelimiter", "comma")
    if delim_name == "tab":
        field_delimiter = "\t"
    elif delim_name == "comma":
        field_delimiter = ","
    else:
        raise ValueError(f"scala.MatchError: {delim_name}")                         # only comma or tab are supported
    ignore_first_line = params.getBoolean("ignoreFirstLine", True)
    if not params.has("input"):
        print("Use --input to specify file input.", file=out)
        return None
    own = ctx is None
    ctx = ctx or N.Context()
    try:
        input_ds = readCsvFile(ctx, params.get("input"), field_delimiter, ignore_first_line)
        als = ALS(ctx).setIterations(params.getInt("iterations", 10)).setNumFactors(params.getInt("numFactors", 10))
        if params.has("blocks"):
            als.setBlocks(params.getInt("blocks"))
        if params.has("temporaryPath"):
            als.setTemporaryPath(params.get("temporaryPath"))
        parameters = ParameterMap().add(ALS.Lambda, params.getDouble("lambda", 0.9)).add(ALS.Seed, params.getLong("seed", 42))
        als.fit(input_ds, parameters)
        users, items = als.factorsOption
        item_rows = [als_factor_row(t.id, "I", t.factors) for t in items]
        user_rows = [als_factor_row(t.id, "U", t.factors) for t in users]
        if params.has("itemFactors") and params.has("userFactors"):
            _write_text(params.get("itemFactors"), item_rows)
            _write_text(params.get("userFactors"), user_rows)
        else:
            print("Printing results to stdout. Use --itemFactors and --userFactors to specify output locations.", file=out)
            print("==== USER FACTORS ====", file=out)
            for r in user_rows:
                print(r, file=out)
            print("==== ITEM FACTORS ====", file=out)
            for r in item_rows:
                print(r, file=out)
        return als
    finally:
        if own:
            ctx.close()


def svm_impl(argv, ctx: N.Context | None = None, out=sys.stdout):
    """SVMImpl.main(args)."""
    params = ParameterTool.fromArgs(argv)
    path = params.getRequired("training")
    own = ctx is None
    ctx = ctx or N.Context()
    try:
        y, csr, dim = readLibSVM(ctx, path)
        svm = SVM(ctx).setBlocks(params.getInt("blocks", 10)).setIterations(params.getInt("iteration", 10))   # sic: "iteration"
        if params.has("seed"):
            svm.setSeed(params.getLong("seed"))                                     # not a flag of the reference job (its seed is random)
        svm.fit(None, y, parallelism=params.getInt("parallelism", 1), csr=csr, dim=dim)
        weights = svm.weightsOption[0].data
        rows = svm_model_rows(weights, partition=params.getBoolean("partition", False), range_=params.getInt("range", 1000))
        if params.has("output"):
            _write_text(params.get("output"), rows)
        else:
            print("Printing result to stdout. Use --output to specify output path.", file=out)
            for r in rows:
                print(r, file=out)
        return svm
    finally:
        if own:
            ctx.close()


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] not in ("als", "svm"):
        print("usage: python -m flinkms_b200 als --input ratings.csv [--fieldDelimiter comma|tab] [--ignoreFirstLine true] "
              "[--iterations 10] [--numFactors 10] [--lambda 0.9] [--seed 42] [--userFactors U.txt --itemFactors I.txt]\n"
              "       python -m flinkms_b200 svm --training data.libsvm [--blocks 10] [--iteration 10] [--partition true --range 1000] "
              "[--output model.txt]", file=sys.stderr)
        return 2
    (als_impl if argv[0] == "als" else svm_impl)(argv[1:])
    return 0
