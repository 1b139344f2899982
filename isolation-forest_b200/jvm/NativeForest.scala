// NOT compiled in this repository's image (no Scala / Spark toolchain).  Drop-in glue for linkedin/isolation-forest:
// the reference keeps its public classes, params and on-disk format; only the three hot bodies call into
// libifb200.so (include/ifb200.h) through the JNI functions of ifb200_jni.c.
package com.linkedin.relevance.isolationforest.gpu

import java.nio.{ByteBuffer, ByteOrder}

import com.linkedin.relevance.isolationforest.IsolationTree
import com.linkedin.relevance.isolationforest.Nodes.{ExternalNode, InternalNode, Node}
import org.apache.spark.ml.linalg.Vector

/** Pre-order node rows of a forest: the layout of NodeData (IsolationForestModelReadWrite.scala:60-67). */
final case class ForestTables(
  nodeOff: Array[Int], left: Array[Int], right: Array[Int], feature: Array[Int],
  threshold: Array[Double], numInstances: Array[Long])

object ForestTables {
  /** Same pre-order numbering as NodeData.build (IsolationForestModelReadWrite.scala:82-132). */
  def fromTrees(trees: Array[IsolationTree]): ForestTables = {
    import scala.collection.mutable.ArrayBuffer
    val off = ArrayBuffer(0)
    val l = ArrayBuffer.empty[Int]; val r = ArrayBuffer.empty[Int]; val f = ArrayBuffer.empty[Int]
    val t = ArrayBuffer.empty[Double]; val n = ArrayBuffer.empty[Long]
    trees.foreach { tree =>
      val base = l.length
      def visit(node: Node): Int = node match { // returns the tree-local id of `node`
        case ExternalNode(numInstances) =>
          val id = l.length - base
          l += -1; r += -1; f += -1; t += 0.0; n += numInstances
          id
        case InternalNode(leftChild, rightChild, splitAttribute, splitValue) =>
          val id = l.length - base
          l += (id + 1); r += -1; f += splitAttribute; t += splitValue; n += -1L
          visit(leftChild)
          r(base + id) = visit(rightChild)
          id
      }
      visit(tree.node)
      off += l.length
    }
    ForestTables(off.toArray, l.toArray, r.toArray, f.toArray, t.toArray, n.toArray)
  }
}

private[isolationforest] object NativeForest {
  System.loadLibrary("ifb200_jni") // links libifb200.so

  @native def hostAlloc(bytes: Long): ByteBuffer
  @native def hostFree(buf: ByteBuffer): Unit
  @native def createStandard(device: Int, nodeOff: Array[Int], left: Array[Int], right: Array[Int],
    feature: Array[Int], threshold: Array[Double], numInstances: Array[Long],
    numSamples: Int, totalNumFeatures: Int): Long
  @native def createExtended(device: Int, nodeOff: Array[Int], left: Array[Int], right: Array[Int],
    numInstances: Array[Long], offset: Array[Double], hpOff: Array[Long], hpIdx: Array[Int], hpW: Array[Float],
    numSamples: Int, totalNumFeatures: Int): Long
  @native def destroy(handle: Long): Unit
  @native def scoreHost(handle: Long, x: ByteBuffer, nRows: Long, d: Int, ld: Long, layout: Int,
    scores: ByteBuffer): Unit
  @native def fitHost(device: Int, x: ByteBuffer, nRows: Long, d: Int, ld: Long, layout: Int,
    numEstimators: Int, numSamples: Int, numFeatures: Int, bootstrap: Boolean, randomSeed: Long,
    numPartitions: Int, extensionLevel: Int, treeBegin: Int, treeEnd: Int): Long

  val RowMajor = 1

  /** One batch of rows of a partition: `.toFloat` into a pinned direct buffer (IsolationForestModel.scala:136),
   *  one native call, scores read back.  Replaces the per-row UDF body (IsolationForestModel.scala:131-139). */
  def scoreBatch(handle: Long, batch: IndexedSeq[Vector], d: Int, x: ByteBuffer, scores: ByteBuffer): Array[Double] = {
    val fb = x.order(ByteOrder.nativeOrder()).asFloatBuffer()
    var i = 0
    while (i < batch.length) {
      val v = batch(i); require(v.size == d,
        s"Input feature vector size ${v.size} did not match the model's training dimension $d.")
      var c = 0
      while (c < d) { fb.put(i * d + c, v(c).toFloat); c += 1 }
      i += 1
    }
    scoreHost(handle, x, batch.length, d, d, RowMajor, scores)
    val db = scores.order(ByteOrder.nativeOrder()).asDoubleBuffer()
    Array.tabulate(batch.length)(db.get)
  }
}
