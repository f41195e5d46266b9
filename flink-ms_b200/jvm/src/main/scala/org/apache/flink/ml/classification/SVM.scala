package org.apache.flink.ml.classification

import java.nio.{ByteBuffer, ByteOrder}

import de.tub.it4bi.gpu.Native
import org.apache.flink.api.common.functions.RichMapPartitionFunction
import org.apache.flink.api.scala._
import org.apache.flink.ml.common.LabeledVector
import org.apache.flink.ml.math.DenseVector
import org.apache.flink.util.Collector

/** Drop-in for org.apache.flink.ml.classification.SVM (FlinkML 1.3.1) as used by flink-svm/SVMImpl.scala:24-34.
  * Block membership mirrors FlinkMLTools.block: blockID = index of the SOURCE subtask % blocks, recorded by a
  * first map stage running at the input's parallelism, before one GPU task gathers everything.
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
    val tagged: DataSet[(Int, LabeledVector)] = input.mapPartition(new RichMapPartitionFunction[LabeledVector, (Int, LabeledVector)] {
      override def mapPartition(it: java.lang.Iterable[LabeledVector], out: Collector[(Int, LabeledVector)]): Unit = {
        val b = getRuntimeContext.getIndexOfThisSubtask % k
        val i = it.iterator(); while (i.hasNext) out.collect((b, i.next()))
      }
    })
    val (iters, h, lam, step, sd) = (iterations, localIterations, regularization, stepsize, seed)
    weightsOption = Some(tagged.mapPartition { it =>
      val rows = it.toArray
      val n = rows.length
      val dim = rows.head._2.vector.size
      def direct(bytes: Long) = ByteBuffer.allocateDirect(bytes.toInt).order(ByteOrder.nativeOrder())
      val (x, y, bid, w) = (direct(8L * n * dim), direct(8L * n), direct(4L * n), direct(8L * dim))
      rows.foreach { case (b, lv) =>
        var j = 0; while (j < dim) { x.putDouble(lv.vector(j)); j += 1 }
        y.putDouble(lv.label); bid.putInt(b)
      }
      val gpu = new Native(0)
      try gpu.svmFit(x, 1, null, null, y, n, dim, bid, k, iters, h, lam, step, sd, 0, w) finally gpu.close()
      Iterator(DenseVector(Array.tabulate(dim)(j => w.getDouble(8 * j))))
    }.setParallelism(1))
  }
}

object SVM { def apply(): SVM = new SVM() }
