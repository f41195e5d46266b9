package org.apache.flink.ml.classification

import java.nio.{ByteBuffer, ByteOrder}

import de.tub.it4bi.gpu.Native
import org.apache.flink.api.scala._
import org.apache.flink.ml.common.LabeledVector
import org.apache.flink.ml.math.DenseVector

/** Drop-in for org.apache.flink.ml.classification.SVM (FlinkML 1.3.1) as used by flink-svm/SVMImpl.scala:24-34.
  * Block membership is FlinkMLTools.block's: blockID = element.hashCode() % blocks, made non-negative (the LabeledVector case
  * class hash over label and vector; flink-ms_b200/flinkml_hash.py restates it for the Python mirror), computed on the JVM where
  * the objects live.  One GPU task then streams the rows to the device in pieces (chunked upload, fp32 on the wire).
  * NOT COMPILED HERE (no JDK/Scala in the build image) -- see INTEGRATION.md. */
class SVM {
  private var blocks: Option[Int] = None
  private var iterations = 10
  private var localIterations = 10
  private var regularization = 1.0
  private var stepsize = 1.0
  private var seed: Long = scala.util.Random.nextLong()
  var weightsOption: Option[DataSet[DenseVector]] = None

  def setBlocks(b: Int): SVM = { blocks = Some(b); this }
  def setIterations(i: Int): SVM = { iterations = i; this }
  def setLocalIterations(i: Int): SVM = { localIterations = i; this }
  def setRegularization(r: Double): SVM = { regularization = r; this }
  def setStepsize(s: Double): SVM = { stepsize = s; this }
  def setSeed(s: Long): SVM = { seed = s; this }

  def fit(input: DataSet[LabeledVector]): Unit = {
    val k = blocks.getOrElse(input.getParallelism)
    val parallelism = input.getParallelism
    val (iters, h, lam, step, sd) = (iterations, localIterations, regularization, stepsize, seed)
    weightsOption = Some(input.mapPartition { it =>
      val rows = it.toArray
      val n = rows.length
      val dim = rows.head.vector.size
      val CHUNK = math.max(1, (1 << 28) / (4 * dim))                        // 256 MiB of fp32 rows per piece
      def direct(bytes: Long) = { require(bytes < Int.MaxValue, "direct buffer over 2 GiB"); ByteBuffer.allocateDirect(bytes.toInt).order(ByteOrder.nativeOrder()) }
      val (x, y, bid, w) = (direct(4L * CHUNK * dim), direct(8L * CHUNK), direct(4L * CHUNK), direct(8L * dim))
      // localDualMethod keeps ONE java.util.Random(seed ^ subtaskIndex) per subtask: parallelism 1 => all blocks share the
      // stream (rngMode 1); parallelism >= blocks => one stream per block (rngMode 0)
      val rngMode = if (parallelism <= 1 && k > 1) 1 else 0
      val gpu = Native.fromProperties()
      try {
        gpu.svmStreamBegin(n, dim)
        var b = 0
        while (b < n) {
          val e = math.min(n, b + CHUNK)
          x.clear(); y.clear(); bid.clear()
          var r = b
          while (r < e) {
            val lv = rows(r)
            var j = 0; while (j < dim) { x.putFloat(lv.vector(j).toFloat); j += 1 }
            y.putDouble(lv.label)
            val hb = lv.hashCode() % k; bid.putInt(if (hb < 0) hb + k else hb)   // FlinkMLTools.block
            r += 1
          }
          gpu.svmStreamAppend(x, 0, y, bid, e - b)
          b = e
        }
        gpu.svmStreamFit(k, iters, h, lam, step, sd, rngMode, w)
      } finally gpu.close()
      Iterator(DenseVector(Array.tabulate(dim)(j => w.getDouble(8 * j))))
    }.setParallelism(1))
  }
}

object SVM { def apply(): SVM = new SVM() }
