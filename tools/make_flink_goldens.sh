#!/bin/bash
# Produces tests/golden/flink_*.json from REAL Flink 1.3.1 + flink-ml_2.11:1.3.1 runs of the reference jobs, so that
# tests/test_oracle.py::test_flink_goldens_if_present can pin the oracle (today: "parity unpinned", no JVM in the image).
#
# Needs: a JDK 8, Maven (or a populated ~/.m2 with org.apache.flink:*:1.3.1), python3 + numpy, network access to Maven
# Central on first use.  Run from the repo root on any machine that has them:
#
#     REFERENCE=/path/to/flink-ms tools/make_flink_goldens.sh
#
# What it does
#   1. builds flink-als and flink-svm of the reference with their own POMs (fat jars, `mvn -q package`);
#   2. writes the inputs: BASELINE.json configs[0] (10k x 10k, 100k ratings, rank 8, seed 42, lambda 0.9, 10 iterations)
#      as a CSV, and a 400 x 24 LibSVM toy set for the SVM;
#   3. runs ALSImpl / SVMImpl in a local Flink environment (parallelism 1, `java -cp <fat jar> <main class>`; the jobs call
#      ExecutionEnvironment.getExecutionEnvironment, which is a LocalEnvironment outside a cluster);
#      SVMImpl cannot set the SVM seed (SVMImpl.scala:24-26; FlinkML draws Random.nextLong()), so the SVM golden is produced
#      by tools/FlinkGoldenSVM.scala (written below) which is SVMImpl plus `.setSeed(42L)` and both LocalIterations 10 and n_k;
#   4. converts the text outputs (`id,U|I,f1;f2;...` rows, ALSImpl.scala:84; `id,w` rows, SVMImpl.scala:35) into
#      tests/golden/flink_als_cfg0.json and tests/golden/flink_svm_toy.json in the schema the test reads.
# The block assignment of every SVM row (which FlinkML derives from LabeledVector.hashCode -- or from the source subtask,
# the open question of the risk register) is dumped by the Scala helper too, so the golden pins that switch as well.
set -euo pipefail
REF=${REFERENCE:?set REFERENCE to a checkout of mmziyad/flink-ms}
OUT=$(mktemp -d)
command -v java >/dev/null || { echo "no JVM on PATH: cannot generate Flink goldens" >&2; exit 3; }
command -v mvn >/dev/null || { echo "no Maven on PATH: cannot build the reference jobs" >&2; exit 3; }

(cd "$REF/flink-als" && mvn -q -DskipTests package)
(cd "$REF/flink-svm" && mvn -q -DskipTests package)
ALS_JAR=$(ls "$REF"/flink-als/target/*.jar | grep -v original | head -1)
SVM_JAR=$(ls "$REF"/flink-svm/target/*.jar | grep -v original | head -1)

python3 - "$OUT" <<'PY'
import importlib, sys, numpy as np
sys.path.insert(0, ".")
synth = importlib.import_module("flink-ms_b200.synth")
out = sys.argv[1]
u, i, r = synth.als_ratings(10_000, 10_000, 100_000, seed=1234)
with open(f"{out}/ratings.csv", "w") as f:
    f.write("user,item,rating\n")
    for a, b, c in zip(u, i, r):
        f.write(f"{a},{b},{c!r}\n")
np.savez(f"{out}/als_in.npz", user=u, item=i, rating=r)
X, y, _ = synth.svm_data(400, 24, seed=5)
with open(f"{out}/toy.libsvm", "w") as f:
    for row, lab in zip(X, y):
        f.write(f"{int(lab)} " + " ".join(f"{k + 1}:{float(v)!r}" for k, v in enumerate(row)) + "\n")
np.savez(f"{out}/svm_in.npz", X=X.astype(np.float64), y=y)
PY

java -cp "$ALS_JAR" de.tub.it4bi.ALSImpl --input "$OUT/ratings.csv" --iterations 10 --numFactors 8 --lambda 0.9 --seed 42 \
     --userFactors "$OUT/users.txt" --itemFactors "$OUT/items.txt"

cat > "$OUT/FlinkGoldenSVM.scala" <<'SCALA'
import org.apache.flink.api.scala._
import org.apache.flink.ml._
import org.apache.flink.ml.classification.SVM
import org.apache.flink.ml.common.FlinkMLTools
object FlinkGoldenSVM {
  def main(args: Array[String]): Unit = {
    val env = ExecutionEnvironment.getExecutionEnvironment
    env.setParallelism(1)
    val data = env.readLibSVM(args(0))
    for (h <- Seq(10, -1)) {
      val svm = SVM().setBlocks(4).setIterations(3).setSeed(42L).setRegularization(1.0).setStepsize(1.0)
      if (h > 0) svm.setLocalIterations(h) else svm.setLocalIterations(100)       // 400 rows / 4 blocks
      svm.fit(data)
      svm.weightsOption.get.map(w => w.data.mkString(",")).writeAsText(args(1) + s"/weights_h$h.txt")
    }
    FlinkMLTools.block(data, 4, None).flatMap(b => b.values.map(v => (b.index, v.label, v.vector.hashCode)))
      .writeAsCsv(args(1) + "/blocks.csv")
    env.execute("golden")
  }
}
SCALA
command -v scalac >/dev/null || { echo "scalac 2.11 needed for the seeded SVM golden; ALS golden is in $OUT" >&2; exit 4; }
scalac -cp "$SVM_JAR" -d "$OUT/golden_svm" "$OUT/FlinkGoldenSVM.scala"
java -cp "$SVM_JAR:$OUT/golden_svm" FlinkGoldenSVM "$OUT/toy.libsvm" "$OUT"

python3 - "$OUT" <<'PY'
import json, sys, numpy as np
out = sys.argv[1]
a = np.load(f"{out}/als_in.npz")
def rows(path):
    d = {}
    for line in open(path):
        k, _, v = line.strip().split(",", 2)
        d[k] = [float(x) for x in v.split(";")]
    return d
json.dump({"kind": "als", "user": a["user"].tolist(), "item": a["item"].tolist(), "rating": a["rating"].tolist(), "rank": 8,
           "iterations": 10, "lambda": 0.9, "seed": 42, "user_factors": rows(f"{out}/users.txt"), "item_factors": rows(f"{out}/items.txt")},
          open("tests/golden/flink_als_cfg0.json", "w"))
s = np.load(f"{out}/svm_in.npz")
blocks = {}
for line in open(f"{out}/blocks.csv"):
    b, lab, h = line.strip().split(",")
    blocks.setdefault(int(h), []).append(int(b))
# row -> block through the dumped (block, label, vector hash) triples, matched by recomputing the hash here
import importlib
sys.path.insert(0, ".")
H = importlib.import_module("flink-ms_b200.flinkml_hash")
vh = H.dense_vector_hash(s["X"]).view(np.int32)
bid = [blocks[int(h)][0] for h in vh]
for h, name in ((10, "flink_svm_toy_h10.json"), (100, "flink_svm_toy_epoch.json")):
    w = [float(x) for x in open(f"{out}/weights_h{h if h == 10 else -1}.txt").read().strip().split(",")]
    json.dump({"kind": "svm", "X": s["X"].tolist(), "y": s["y"].tolist(), "block_id": bid, "blocks": 4, "iterations": 3,
               "local_iterations": h, "regularization": 1.0, "stepsize": 1.0, "seed": 42, "rng_mode": 1, "weights": w},
              open(f"tests/golden/{name}", "w"))
PY
echo "goldens written to tests/golden/flink_*.json -- run: python -m pytest tests/test_oracle.py -k flink_goldens"
