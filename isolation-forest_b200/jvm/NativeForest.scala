// NOT compiled in this repository's image (no Scala / Spark toolchain); tests/test_jni_glue.py checks that every @native
// method below has its C implementation in ifb200_jni.c (and vice versa).  Drop-in glue for linkedin/isolation-forest:
// the reference keeps its public classes, params and on-disk format; only the three hot bodies call into
// libifb200.so (include/ifb200.h) through the JNI functions of ifb200_jni.c.
package com.linkedin.relevance.isolationforest.gpu

import java.nio.{ByteBuffer, ByteOrder}

import com.linkedin.relevance.isolationforest.IsolationTree
import com.linkedin.relevance.isolationforest.Nodes.{ExternalNode, InternalNode, Node}
import com.linkedin.relevance.isolationforest.extended.ExtendedIsolationTree
import com.linkedin.relevance.isolationforest.extended.ExtendedNodes.{ExtendedExternalNode, ExtendedInternalNode, ExtendedNode}
import com.linkedin.relevance.isolationforest.extended.ExtendedUtils.SplitHyperplane
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

/** Pre-order node rows of an extended forest: the layout of ExtendedNodeData
 *  (extended/ExtendedIsolationForestModelReadWrite.scala:59-67); hyperplanes are CSR over the node rows. */
final case class ExtendedForestTables(
  nodeOff: Array[Int], left: Array[Int], right: Array[Int], numInstances: Array[Long], offset: Array[Double],
  hpOff: Array[Long], hpIdx: Array[Int], hpW: Array[Float])

object ExtendedForestTables {
  def fromTrees(trees: Array[ExtendedIsolationTree]): ExtendedForestTables = {
    import scala.collection.mutable.ArrayBuffer
    val off = ArrayBuffer(0)
    val l = ArrayBuffer.empty[Int]; val r = ArrayBuffer.empty[Int]; val n = ArrayBuffer.empty[Long]
    val o = ArrayBuffer.empty[Double]; val ho = ArrayBuffer(0L)
    val hi = ArrayBuffer.empty[Int]; val hw = ArrayBuffer.empty[Float]
    trees.foreach { tree =>
      val base = l.length
      def visit(node: ExtendedNode): Int = node match {
        case ExtendedExternalNode(numInstances) =>
          val id = l.length - base
          l += -1; r += -1; n += numInstances; o += 0.0; ho += hi.length.toLong
          id
        case ExtendedInternalNode(leftChild, rightChild, hp) =>
          val id = l.length - base
          l += (id + 1); r += -1; n += -1L; o += hp.offset
          hi ++= hp.indices; hw ++= hp.weights; ho += hi.length.toLong
          visit(leftChild)
          r(base + id) = visit(rightChild)
          id
      }
      visit(tree.extendedNode)
      off += l.length
    }
    ExtendedForestTables(off.toArray, l.toArray, r.toArray, n.toArray, o.toArray, ho.toArray, hi.toArray, hw.toArray)
  }

  /** Inverse of fromTrees: what buildTreeFromNodes does with the Avro rows
   *  (extended/ExtendedIsolationForestModelReadWrite.scala:179-211). */
  def toTrees(t: ExtendedForestTables): Array[ExtendedIsolationTree] =
    Array.tabulate(t.nodeOff.length - 1) { k =>
      val base = t.nodeOff(k)
      def build(id: Int): ExtendedNode = {
        val g = base + id
        if (t.left(g) == -1) ExtendedExternalNode(t.numInstances(g))
        else {
          val (b, e) = (t.hpOff(g).toInt, t.hpOff(g + 1).toInt)
          ExtendedInternalNode(build(t.left(g)), build(t.right(g)),
            SplitHyperplane(t.hpIdx.slice(b, e), t.hpW.slice(b, e), t.offset(g)))
        }
      }
      new ExtendedIsolationTree(build(0))
    }
}

object ForestTablesOps {
  /** Inverse of ForestTables.fromTrees (IsolationForestModelReadWrite.scala:179-205 buildTreeFromNodes). */
  def toTrees(t: ForestTables): Array[IsolationTree] =
    Array.tabulate(t.nodeOff.length - 1) { k =>
      val base = t.nodeOff(k)
      def build(id: Int): Node = {
        val g = base + id
        if (t.left(g) == -1) ExternalNode(t.numInstances(g))
        else InternalNode(build(t.left(g)), build(t.right(g)), t.feature(g), t.threshold(g))
      }
      new IsolationTree(build(0))
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
  /** {extended, device, numTrees, numSamples, totalNumFeatures, maxFeatureIndex, maxDepth, maxNnz, numNodes, numHpEntries} */
  @native def info(handle: Long): Array[Long]
  /** Node tables of a (GPU-fitted) forest in the persisted layout, see ifb200_jni.c exportTables. */
  @native def exportTables(handle: Long): Array[AnyRef]
  @native def scoreHost(handle: Long, x: ByteBuffer, nRows: Long, d: Int, ld: Long, layout: Int,
    scores: ByteBuffer): Unit
  @native def fitHost(device: Int, x: ByteBuffer, nRows: Long, d: Int, ld: Long, layout: Int,
    numEstimators: Int, numSamples: Int, numFeatures: Int, bootstrap: Boolean, randomSeed: Long,
    numPartitions: Int, extensionLevel: Int, treeBegin: Int, treeEnd: Int): Long

  @native def quantileHost(device: Int, scores: ByteBuffer, n: Long, q: Double): Array[Double]
  @native def commUniqueId(): Array[Byte]
  @native def commInit(device: Int, world: Int, rank: Int, id: Array[Byte]): Long
  @native def commDestroy(comm: Long): Unit
  @native def scoreShardedHost(handle: Long, comm: Long, device: Int, x: ByteBuffer, nRows: Long, d: Int,
    totalNumTrees: Int, scores: ByteBuffer): Unit

  val RowMajor = 1

  def create(device: Int, t: ForestTables, numSamples: Int, totalNumFeatures: Int): Long =
    createStandard(device, t.nodeOff, t.left, t.right, t.feature, t.threshold, t.numInstances, numSamples, totalNumFeatures)

  def create(device: Int, t: ExtendedForestTables, numSamples: Int, totalNumFeatures: Int): Long =
    createExtended(device, t.nodeOff, t.left, t.right, t.numInstances, t.offset, t.hpOff, t.hpIdx, t.hpW, numSamples,
      totalNumFeatures)

  /** Trees of a forest fitted on the GPU (fitHost), ready for IsolationForestModel / the Avro writer. */
  def standardTrees(handle: Long): Array[IsolationTree] = {
    val a = exportTables(handle)
    ForestTablesOps.toTrees(ForestTables(a(0).asInstanceOf[Array[Int]], a(1).asInstanceOf[Array[Int]],
      a(2).asInstanceOf[Array[Int]], a(3).asInstanceOf[Array[Int]], a(4).asInstanceOf[Array[Double]],
      a(5).asInstanceOf[Array[Long]]))
  }

  def extendedTrees(handle: Long): Array[ExtendedIsolationTree] = {
    val a = exportTables(handle)
    ExtendedForestTables.toTrees(ExtendedForestTables(a(0).asInstanceOf[Array[Int]], a(1).asInstanceOf[Array[Int]],
      a(2).asInstanceOf[Array[Int]], a(3).asInstanceOf[Array[Long]], a(4).asInstanceOf[Array[Double]],
      a(5).asInstanceOf[Array[Long]], a(6).asInstanceOf[Array[Int]], a(7).asInstanceOf[Array[Float]]))
  }

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
