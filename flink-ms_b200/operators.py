"""Host-side mirror of the operator surface the reference programs against:

    ALS().setIterations(..).setNumFactors(..).setBlocks(..).setTemporaryPath(..)      ALSImpl.scala:35-44
    ParameterMap().add(ALS.Lambda, 0.9).add(ALS.Seed, 42L); als.fit(ds, params)       ALSImpl.scala:47-52
    als.factorsOption -> (userFactors, itemFactors) of Factors(id, factors)           ALSImpl.scala:55-57
    SVM().setBlocks(..).setIterations(..); svm.fit(ds); svm.weightsOption             SVMImpl.scala:24-34

Same names, same argument meaning, same defaults (FlinkML 1.3.1 defaults where the reference does not
override them), same error behaviour (predict/factors before fit raises).  All arithmetic happens in
libflinkms_b200.so through the C-ABI; there is no CPU path here.  One difference is deliberate and
documented: FlinkML's fit() is lazy (it only builds a DataSet plan, run by env.execute()); this mirror
runs eagerly -- the JVM facade in INTEGRATION.md keeps the laziness.
"""
from __future__ import annotations

import random
from dataclasses import dataclass

import numpy as np

from . import _native as N
from . import flinkml_hash as H


class ParameterMap:
    """org.apache.flink.ml.common.ParameterMap: `.add(key, value)` is fluent (ALSImpl.scala:47-49)."""

    def __init__(self):
        self._m = {}

    def add(self, key, value):
        self._m[key] = value
        return self

    def get(self, key, default=None):
        return self._m.get(key, default)

    def __contains__(self, key):
        return key in self._m


@dataclass(frozen=True)
class Factors:
    """org.apache.flink.ml.recommendation.ALS.Factors(id: Long, factors: Array[Double]) (ALSImpl.scala:55-57, :83)."""
    id: int
    factors: np.ndarray


class _FactorSet:
    """Stands in for DataSet[Factors]: iterable of Factors plus the packed arrays."""

    def __init__(self, ids, factors):
        self.ids = ids
        self.factors = factors

    def __len__(self):
        return len(self.ids)

    def __iter__(self):
        for k in range(len(self.ids)):
            yield Factors(int(self.ids[k]), self.factors[k])

    def collect(self):
        return list(self)


class ALS:
    """org.apache.flink.ml.recommendation.ALS (FlinkML 1.3.1) on B200."""

    # ParameterMap keys (ALS.Lambda / ALS.Seed in ALSImpl.scala:48-49) and their FlinkML defaults
    NumFactors = "ALS.NumFactors"
    Lambda = "ALS.Lambda"
    Iterations = "ALS.Iterations"
    Blocks = "ALS.Blocks"
    Seed = "ALS.Seed"
    TemporaryPath = "ALS.TemporaryPath"
    _DEFAULTS = {NumFactors: 10, Lambda: 1.0, Iterations: 10, Blocks: None, Seed: 0, TemporaryPath: None}

    def __init__(self, ctx: N.Context | None = None):
        self.parameters = ParameterMap()
        self._ctx = ctx
        self._own_ctx = ctx is None
        self.factorsOption = None          # Option[(DataSet[Factors], DataSet[Factors])] = (users, items)

    # fluent setters -------------------------------------------------------------------------------------
    def setIterations(self, iterations: int):
        self.parameters.add(ALS.Iterations, int(iterations)); return self

    def setNumFactors(self, numFactors: int):
        self.parameters.add(ALS.NumFactors, int(numFactors)); return self

    def setLambda(self, lam: float):
        self.parameters.add(ALS.Lambda, float(lam)); return self

    def setBlocks(self, blocks: int):
        # block count only changes the fp64 summation order in Flink; the GPU path has no blocks
        self.parameters.add(ALS.Blocks, int(blocks)); return self

    def setSeed(self, seed: int):
        self.parameters.add(ALS.Seed, int(seed)); return self

    def setTemporaryPath(self, path: str):
        # Flink persists intermediate block data there to cut the plan; nothing to persist here
        self.parameters.add(ALS.TemporaryPath, path); return self

    def _param(self, key, fit_parameters):
        if fit_parameters is not None and key in fit_parameters:
            return fit_parameters.get(key)
        if key in self.parameters:
            return self.parameters.get(key)
        return ALS._DEFAULTS[key]

    def _context(self):
        if self._ctx is None:
            self._ctx = N.Context()
        return self._ctx

    # fit ------------------------------------------------------------------------------------------------
    def fit(self, ratings, fitParameters: ParameterMap | None = None):
        """ratings: (user[], item[], rating[]) arrays, or an iterable of (user, item, rating) triples
        (the DataSet[(Int, Int, Double)] of ALSImpl.scala:29-32)."""
        if isinstance(ratings, tuple) and len(ratings) == 3 and not np.isscalar(ratings[0]):
            user, item, rating = ratings
        else:
            arr = list(ratings)
            user = [t[0] for t in arr]; item = [t[1] for t in arr]; rating = [t[2] for t in arr]
        rank = int(self._param(ALS.NumFactors, fitParameters))
        iterations = int(self._param(ALS.Iterations, fitParameters))
        lam = float(self._param(ALS.Lambda, fitParameters))
        seed = int(self._param(ALS.Seed, fitParameters))
        ctx = self._context()
        ctx.als_fit(user, item, rating, rank, iterations, lam, seed)
        uid, U = ctx.als_get_factors(N.USERS)
        iid, I = ctx.als_get_factors(N.ITEMS)
        self.factorsOption = (_FactorSet(uid, U), _FactorSet(iid, I))

    def predict(self, pairs):
        """ALS.predict: (user, item) pairs -> (user, item, prediction) for pairs whose ids are both in the
        model (inner join: unknown ids are dropped).  Returns (users, items, predictions) arrays."""
        if self.factorsOption is None:
            raise RuntimeError("The ALS model has not been fitted. Prior to predicting values, it has to be "
                               "trained on data.")
        if isinstance(pairs, tuple) and len(pairs) == 2 and not np.isscalar(pairs[0]):
            u, i = np.asarray(pairs[0], np.int64), np.asarray(pairs[1], np.int64)
        else:
            arr = list(pairs)
            u = np.asarray([t[0] for t in arr], np.int64); i = np.asarray([t[1] for t in arr], np.int64)
        out, found = self._context().als_predict(u, i)
        return u[found], i[found], out[found]

    def empiricalRisk(self, ratings):
        raise NotImplementedError("ALS.empiricalRisk is never called by the reference (SURVEY.md Appendix A.1)")


def readCsvFile(ctx: N.Context, filePath: str, fieldDelimiter: str = ",", ignoreFirstLine: bool = False):
    """env.readCsvFile[(Int, Int, Double)](filePath, fieldDelimiter = ..., ignoreFirstLine = ...) of ALSImpl.scala:29-32,
    parsed on the GPU (row N4): returns the (user[], item[], rating[]) arrays ALS.fit takes."""
    with open(filePath, "rb") as f:
        return ctx.parse_ratings_csv(f.read(), fieldDelimiter, ignoreFirstLine)


def readLibSVM(ctx: N.Context, filePath: str):
    """env.readLibSVM(filePath) of SVMImpl.scala:21 (FlinkML MLUtils.readLibSVM), parsed on the GPU (row N4): returns
    (labels, (csr_ptr, csr_idx, csr_val), dimension) -- the sparse LabeledVector data set as CSR."""
    with open(filePath, "rb") as f:
        return ctx.parse_libsvm(f.read())


class OnlineSGD:
    """The per-rating update of the als-ms `SGD` job (als-ms/.../qs/SGD.java:160-217) as a micro-batch operator over a
    GPU-resident model.  Parameter names are the job's CLI flags (`learningRate`, `userRegularization`,
    `itemRegularization`, `userMean`, `itemMean` as "f1;f2;..." strings, SGD.java:143-147, :160-166).  `step` returns the
    records the reference would send to Kafka: "id,U,f1;f2;..." then "id,I,f1;f2;..." per rating (SGD.java:209-216)."""

    def __init__(self, ctx: N.Context, params: dict | None = None):
        self._ctx = ctx
        p = dict(params or {})
        self.learningRate = float(p.get("learningRate", 0.1))
        self.userRegularization = float(p.get("userRegularization", 0.0))
        self.itemRegularization = float(p.get("itemRegularization", 0.0))
        self.userMean = self._vec(p.get("userMean"))
        self.itemMean = self._vec(p.get("itemMean"))

    @staticmethod
    def _vec(v):
        if v is None:
            return None
        if isinstance(v, str):
            return np.array([float(t) for t in v.split(";")], np.float64)
        return np.asarray(v, np.float64)

    def step(self, ratings, apply: bool = True):
        """ratings: (user[], item[], rating[]) arrays or an iterable of (user, item, rating) triples."""
        from .formats import als_factor_row
        if isinstance(ratings, tuple) and len(ratings) == 3 and not np.isscalar(ratings[0]):
            user, item, rating = ratings
        else:
            arr = list(ratings)
            user = [t[0] for t in arr]; item = [t[1] for t in arr]; rating = [t[2] for t in arr]
        if self.userMean is None or self.itemMean is None:
            raise RuntimeError("Unable to load User mean or item mean factors.")   # IllegalStateException, SGD.java:150-152
        ou, oi = self._ctx.als_sgd_step(user, item, rating, self.learningRate, self.userRegularization, self.itemRegularization,
                                        self.userMean, self.itemMean, apply)
        out = []
        for k in range(len(ou)):
            out.append(als_factor_row(user[k], "U", ou[k]))
            out.append(als_factor_row(item[k], "I", oi[k]))
        return out


class DenseVector:
    """org.apache.flink.ml.math.DenseVector: `.data` is the Array[Double] (SVMImpl.scala:34; SGDV0.java:258)."""

    def __init__(self, data):
        self.data = np.asarray(data, np.float64)

    def size(self):
        return len(self.data)

    def apply(self, i):
        return float(self.data[i])

    def update(self, i, v):
        self.data[i] = v

    def dot(self, other):
        return float(np.dot(self.data, other.data if isinstance(other, DenseVector) else np.asarray(other)))


class SVM:
    """org.apache.flink.ml.classification.SVM (FlinkML 1.3.1, CoCoA + SDCA) on B200."""

    _DEFAULTS = dict(blocks=None, iterations=10, local_iterations=10, regularization=1.0, stepsize=1.0, seed=None,
                     threshold=0.0, output_decision_function=False)

    def __init__(self, ctx: N.Context | None = None):
        self._p = dict(SVM._DEFAULTS)
        self._ctx = ctx
        self.weightsOption = None          # Option[DataSet[DenseVector]] (one element)

    def setBlocks(self, blocks: int):
        self._p["blocks"] = int(blocks); return self

    def setIterations(self, iterations: int):
        self._p["iterations"] = int(iterations); return self

    def setLocalIterations(self, localIterations: int):
        self._p["local_iterations"] = int(localIterations); return self

    def setRegularization(self, regularization: float):
        self._p["regularization"] = float(regularization); return self

    def setStepsize(self, stepsize: float):
        self._p["stepsize"] = float(stepsize); return self

    def setSeed(self, seed: int):
        self._p["seed"] = int(seed); return self

    def setThreshold(self, threshold: float):
        self._p["threshold"] = float(threshold); return self

    def setOutputDecisionFunction(self, outputDecisionFunction: bool):
        self._p["output_decision_function"] = bool(outputDecisionFunction); return self

    def _context(self):
        if self._ctx is None:
            self._ctx = N.Context()
        return self._ctx

    @staticmethod
    def default_block_ids(n: int, blocks: int, parallelism: int = 1) -> np.ndarray:
        """The round-1 reading of FlinkMLTools.block (blockID = sourceSubtaskIndex % blocks); see
        flinkml_hash.block_ids_by_subtask.  `fit` now defaults to the hash reading (block_mode="hash")."""
        return H.block_ids_by_subtask(n, blocks, parallelism)

    def fit(self, X, y=None, block_id=None, parallelism: int = 1, csr=None, dim=None, rng_mode=None, block_mode: str = "hash"):
        """X: dense (n x D) array, or csr=(ptr, idx, val) with dim; y in {+1,-1}.
        block_id: explicit CoCoA block of every row.  Default (block_mode="hash"): FlinkMLTools.block as recollected by
        the round-1 advisor and this author -- LabeledVector.hashCode() % blocks (flinkml_hash.py); block_mode="subtask"
        is the round-1 reading.  rng_mode default: localDualMethod keeps ONE java.util.Random(seed ^ subtaskIndex) per
        parallel subtask and the blocks of a subtask draw from it in ascending block order, so parallelism 1 means the
        shared stream (RNG_SHARED_SUBTASK) and parallelism >= blocks one stream per block (RNG_PER_BLOCK); the cases in
        between (several blocks on some subtasks, not all) are not covered by either schedule and are refused."""
        n = len(y)
        blocks = self._p["blocks"] if self._p["blocks"] is not None else parallelism
        if block_id is None:
            if block_mode == "hash":
                block_id = H.block_ids_by_hash(y, blocks, X=X, csr=csr, dim=dim)
            elif block_mode == "subtask":
                block_id = H.block_ids_by_subtask(n, blocks, parallelism)
            else:
                raise ValueError("block_mode must be 'hash' or 'subtask'")
        if rng_mode is None:
            if parallelism <= 1 or blocks == 1:
                rng_mode = N.RNG_SHARED_SUBTASK if blocks > 1 else N.RNG_PER_BLOCK
            elif parallelism >= blocks:
                rng_mode = N.RNG_PER_BLOCK
            else:
                raise NotImplementedError("1 < parallelism < blocks: blocks share per-subtask random streams in a pattern neither "
                                          "RNG schedule of the C-ABI reproduces; pass rng_mode explicitly")
        seed = self._p["seed"]
        if seed is None:
            seed = random.getrandbits(63) - (1 << 62)   # FlinkML default: Random.nextLong()
        ctx = self._context()
        w = ctx.svm_fit(X, y, block_id, blocks, self._p["iterations"], self._p["local_iterations"],
                        self._p["regularization"], self._p["stepsize"], seed, rng_mode, csr=csr, dim=dim)
        self.weightsOption = [DenseVector(w)]

    def predict(self, X, csr=None, dim=None, n=None):
        if self.weightsOption is None:
            raise RuntimeError("The SVM model has not been trained. Call first fit before calling the predict "
                               "operation.")
        w = self.weightsOption[0].data
        return self._context().svm_predict(X, w, self._p["threshold"], self._p["output_decision_function"], csr=csr,
                                           dim=dim if dim is not None else len(w), n=n)
