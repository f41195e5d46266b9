package org.apache.flink.ml.recommendation

import java.nio.{ByteBuffer, ByteOrder}

import de.tub.it4bi.gpu.Native
import org.apache.flink.api.scala._
import org.apache.flink.ml.common.{Parameter, ParameterMap}

/** Drop-in for org.apache.flink.ml.recommendation.ALS (FlinkML 1.3.1) as used by flink-als/ALSImpl.scala:35-57:
  * same class / method names, same parameter keys and defaults, same laziness (fit only appends operators; the GPU
  * work runs inside a parallelism-1 mapPartition when the plan executes), same factorsOption shape.
  * NOT COMPILED HERE (no JDK/Scala in the build image) -- see INTEGRATION.md. */
class ALS {
  val parameters = new ParameterMap()
  var factorsOption: Option[(DataSet[ALS.Factors], DataSet[ALS.Factors])] = None

  def setNumFactors(n: Int): ALS = { parameters.add(ALS.NumFactors, n); this }
  def setLambda(l: Double): ALS = { parameters.add(ALS.Lambda, l); this }
  def setIterations(i: Int): ALS = { parameters.add(ALS.Iterations, i); this }
  def setBlocks(b: Int): ALS = { parameters.add(ALS.Blocks, b); this }          // accepted, no effect on the GPU path
  def setSeed(s: Long): ALS = { parameters.add(ALS.Seed, s); this }
  def setTemporaryPath(p: String): ALS = { parameters.add(ALS.TemporaryPath, p); this }   // accepted, nothing to persist

  /** ALSImpl.scala:52 calls fit(DataSet[(Int, Int, Double)], ParameterMap). */
  def fit(input: DataSet[(Int, Int, Double)], fitParameters: ParameterMap): Unit = {
    val p = parameters ++ fitParameters
    val rank = p(ALS.NumFactors); val iters = p(ALS.Iterations); val lambda = p(ALS.Lambda); val seed = p(ALS.Seed)
    // one GPU task drains the DataSet in pieces of CHUNK triples (a direct ByteBuffer holds < 2 GiB; configs[1] is 2.4 GB)
    // through the chunked-upload entry points, fits, emits (which, id, factors).  Devices: -Dflinkms.devices=0[,1,...]
    val rows: DataSet[(Int, Long, Array[Double])] = input.mapPartition { it =>
      val triples = it.toArray
      val n = triples.length
      val CHUNK = 1 << 24
      def direct(bytes: Long) = { require(bytes < Int.MaxValue, "direct buffer over 2 GiB"); ByteBuffer.allocateDirect(bytes.toInt).order(ByteOrder.nativeOrder()) }
      val (u, i, r) = (direct(8L * CHUNK), direct(8L * CHUNK), direct(8L * CHUNK))
      val gpu = Native.fromProperties()
      try {
        gpu.alsStreamBegin(n)
        var b = 0
        while (b < n) {
          val e = math.min(n, b + CHUNK)
          u.clear(); i.clear(); r.clear()
          var k = b; while (k < e) { val t = triples(k); u.putLong(t._1.toLong); i.putLong(t._2.toLong); r.putDouble(t._3); k += 1 }
          gpu.alsStreamAppend(u, i, r, e - b)
          b = e
        }
        val counts = gpu.alsStreamFit(rank, iters, lambda, seed)
        (0 to 1).iterator.flatMap { which =>
          val cnt = counts(which).toInt
          val ids = direct(8L * cnt); val fac = direct(8L * cnt * rank)
          gpu.alsGetFactors(which, ids, fac)
          (0 until cnt).iterator.map { k =>
            (which, ids.getLong(8 * k), Array.tabulate(rank)(j => fac.getDouble(8 * (k * rank + j))))
          }
        }.toArray.iterator
      } finally gpu.close()
    }.setParallelism(1)
    factorsOption = Some((rows.filter(_._1 == 0).map(t => ALS.Factors(t._2, t._3)),
                          rows.filter(_._1 == 1).map(t => ALS.Factors(t._2, t._3))))
  }
}

object ALS {
  case class Factors(id: Long, factors: Array[Double])                // ALSImpl.scala:55-57, :83
  case object NumFactors extends Parameter[Int] { val defaultValue: Option[Int] = Some(10) }
  case object Lambda extends Parameter[Double] { val defaultValue: Option[Double] = Some(1.0) }
  case object Iterations extends Parameter[Int] { val defaultValue: Option[Int] = Some(10) }
  case object Blocks extends Parameter[Int] { val defaultValue: Option[Int] = None }
  case object Seed extends Parameter[Long] { val defaultValue: Option[Long] = Some(0L) }
  case object TemporaryPath extends Parameter[String] { val defaultValue: Option[String] = None }
  def apply(): ALS = new ALS()
}
