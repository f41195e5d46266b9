"""The two training jobs of the reference as drop-in drivers over the GPU operators.

`als_impl(argv)` mirrors `flink-als/src/main/scala/de/tub/it4bi/ALSImpl.scala:15-75` and `svm_impl(argv)` mirrors
`flink-svm/src/main/scala/de/tub/it4bi/SVMImpl.scala:13-56`: same `--key value` flags (Flink's ParameterTool), same
defaults, same text rows on stdout or in the output files.  The inputs are parsed on the GPU (row N4), the models are
trained by the CUDA kernels behind `operators.ALS` / `operators.SVM`, the rows are formatted like Java does
(`formats.py`).  Run as  `python -m flinkms_b200 als --input ratings.csv --numFactors 64 ...`  (the import shim at the repo root dispatches here).

Differences a user should know: `--blocks` / `--temporaryPath` of the ALS job are accepted and have no effect (no Flink
shuffles here); an output path is written as ONE text file (Flink writes a directory of part files at parallelism > 1);
the SVM block of a row is FlinkMLTools.block's `LabeledVector.hashCode() % blocks` (flinkml_hash.py) and, at the default
`--parallelism 1`, all blocks draw from ONE java.util.Random(seed ^ 0) in ascending block order like the single
localDualMethod subtask of the reference.
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

    _REQUIRED = object()

    def getBoolean(self, key, default=_REQUIRED):
        """getBoolean(key) without a default is getRequired + Boolean.valueOf, like Flink's."""
        if key not in self.data:
            if default is ParameterTool._REQUIRED:
                raise RuntimeError(f"No data for required key '{key}'")
            return default
        return self.data[key].lower() == "true"                                     # Boolean.valueOf


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
        # params.getBoolean("partition") has no default in the reference (SVMImpl.scala:41): the flag is required
        rows = svm_model_rows(weights, partition=params.getBoolean("partition"), range_=params.getInt("range", 1000))
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


def als_mean_vector(argv, out=sys.stdout):
    """ALSMeanVector.main(args) (flink-als/.../ALSMeanVector.scala:12-48): the mean of the factor rows of an ALS output file as
    a `MEAN,U|I,m1;m2;...` row (the cold-start fallback of the online SGD job).  Host arithmetic, like the reference:
    a left-to-right fp64 sum divided by the row count."""
    params = ParameterTool.fromArgs(argv)
    t = params.getRequired("type")
    if t == "item":
        factor_type = "I"
    elif t == "user":
        factor_type = "U"
    else:
        raise ValueError("specify type as either 'item' or 'user'.")                 # IllegalArgumentException
    from .formats import als_mean_row
    with open(params.get("input")) as f:
        vectors = [[float(v) for v in row.rstrip("\n").split(",")[2].split(";")] for row in f if row.strip()]
    row = als_mean_row(np.array(vectors, np.float64), factor_type)
    if params.has("output"):
        _write_text(params.get("output"), [row])
    else:
        print("Printing results to stdout. Use --output to specify output location", file=out)
        print(row, file=out)
    return row


def als_predict_console(lines, client, out=sys.stdout):
    """The query loop of flink-queryable-client/.../qs/ALSPredict.java:58-92 over a `formats.QueryableStateClient`: one
    `user,item` pair per line -> two state lookups ("<user>-U", "<item>-I"), the dot product of the parsed factor strings."""
    print("Enter <User,Item> to predict.", file=out)
    for line in lines:
        key = line.upper().strip()
        out.write("[info] Querying the model for <user,item> pair '%s'\n" % key)
        try:
            tokens = key.split(",")
            user_t = client.queryState("ALS_MODEL", tokens[0] + "-U")
            item_t = client.queryState("ALS_MODEL", tokens[1] + "-I")
            if user_t is not None and item_t is not None:
                uf = [float(t) for t in user_t[1].split(";")]
                itf = [float(t) for t in item_t[1].split(";")]
                if len(uf) != len(itf):
                    raise ValueError("DimensionMismatchException: %d != %d" % (len(itf), len(uf)))
                prediction = 0.0
                for a, b in zip(uf, itf):                                           # ArrayRealVector.dotProduct: sequential
                    prediction += a * b
                out.write("ALS Prediction =  %f \n" % prediction)
            else:
                out.write("User or Item Factors do not exist in the model for the query: ")   # sic: no newline, key dropped by printf
        except Exception as e:  # noqa: BLE001  (the reference catches Exception and prints the stack trace)
            out.write("Query failed because of the following Exception:\n%s: %s\n" % (type(e).__name__, e))


def svm_predict_console(lines, client, outputDecisionFunction=False, thresholdValue=0.0, out=sys.stdout):
    """The query loop of flink-queryable-client/.../qs/SVMPredict.java:58-90: one sparse vector `id:val id:val ...` per line,
    one state lookup per feature id, raw value = sum of weight * value, strict `>` against the threshold."""
    print("Enter Vector data to predict.", file=out)
    for line in lines:
        out.write("[info] Querying the model for vector '%s' \n" % line)
        try:
            raw = 0.0
            for tok in line.strip().split(" "):
                pair = tok.split(":")
                fid, val = pair[0], float(pair[1])
                model_val = client.queryState("SVM_MODEL", fid)
                if model_val is not None:
                    raw += float(model_val[1]) * val
                else:
                    out.write("Could not find the value for feature ID: %s \n" % fid)
            prediction = raw if outputDecisionFunction else (1.0 if raw > thresholdValue else -1.0)
            out.write("SVM Prediction =  %f \n" % prediction)
        except Exception as e:  # noqa: BLE001
            out.write("Query failed because of the following Exception:\n%s: %s\n" % (type(e).__name__, e))


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if argv and argv[0] == "als-mean":
        als_mean_vector(argv[1:])
        return 0
    if not argv or argv[0] not in ("als", "svm"):
        print("usage: python -m flinkms_b200 als --input ratings.csv [--fieldDelimiter comma|tab] [--ignoreFirstLine true] "
              "[--iterations 10] [--numFactors 10] [--lambda 0.9] [--seed 42] [--userFactors U.txt --itemFactors I.txt]\n"
              "       python -m flinkms_b200 svm --training data.libsvm --partition true|false [--range 1000] [--blocks 10] [--iteration 10] "
              "[--output model.txt]\n"
              "       python -m flinkms_b200 als-mean --type user|item --input U.txt [--output mean.txt]", file=sys.stderr)
        return 2
    (als_impl if argv[0] == "als" else svm_impl)(argv[1:])
    return 0
