/*
 * Reference-side binding (NOT compiled in this image: no JVM).  Lives in package com.fulcrumgenomics.identifyprimers so that it can
 * see the private[identifyprimers] types, and replaces the body of IdentifyPrimers.processBatch (IdentifyPrimers.scala:381-436):
 * the per-read Future plumbing becomes ONE native call per 5,000-template batch.
 *
 * `Native` declares every entry of bindings/jni/idprimer_jni.c, i.e. every export of include/idprimer.h; the class below uses the
 * ones the matching path needs.
 */
package com.fulcrumgenomics.identifyprimers

import java.nio.{ByteBuffer, ByteOrder}

import com.fulcrumgenomics.bam.Template
import com.fulcrumgenomics.bam.api.SamRecord

private[identifyprimers] object GpuPrimerMatcher {
  object Native {
    System.loadLibrary("idprimer_jni")
    // library
    @native def abiVersion(): Int
    @native def lastError(): String
    @native def deviceCount(): Int
    @native def alignPolicy(): Int
    // panel: the three matchers' immutable state (IP:223-234)
    @native def panelCreate(primers: ByteBuffer, n: Int, params: ByteBuffer): Long
    @native def panelDestroy(h: Long): Unit
    @native def panelSize(h: Long): Int
    @native def panelPrimerLength(h: Long, i: Int): Int
    @native def panelPrimerHash(h: Long, i: Int): Int
    @native def panelKmerPostings(h: Long, which: Int, kmer: String, out: ByteBuffer, cap: Int): Int
    @native def panelWindow(panel: Long): Int  // idp_panel_window: bases the 5' chain can reach, -1 with --three-prime
    // context: one per worker thread
    @native def ctxCreate(panel: Long, device: Int): Long
    @native def ctxDestroy(h: Long): Unit
    @native def ctxSetStream(h: Long, cudaStream: Long): Unit
    @native def ctxSetAsync(h: Long, enable: Int): Unit
    @native def ctxSync(h: Long, metrics160: ByteBuffer, nAligned: ByteBuffer): Unit
    @native def ctxTimings(h: Long, out: ByteBuffer): Unit
    @native def ctxSetClassify(h: Long, minInsert: Int, maxInsert: Int, clsPerRead: ByteBuffer, counts6: ByteBuffer): Unit
    @native def ctxSetClassifyDevice(h: Long, minInsert: Int, maxInsert: Int, clsPerRead: Long, counts6: ByteBuffer): Unit
    @native def hostAlloc(bytes: Long): ByteBuffer  // pinned
    @native def hostFree(buf: ByteBuffer): Unit
    // processBatch (IP:381-436): 32-byte and 16-byte records, host and device pointers
    @native def matchBatch(ctx: Long, seq4: ByteBuffer, seqOff: ByteBuffer, len: ByteBuffer, refIdx: ByteBuffer, uStart: ByteBuffer,
                           uEnd: ByteBuffer, flags: ByteBuffer, fixedLen: Int, n: Int, five: ByteBuffer, three: ByteBuffer,
                           metrics160: ByteBuffer, nAligned: ByteBuffer): Unit
    @native def matchBatch16(ctx: Long, seq4: ByteBuffer, seqOff: ByteBuffer, len: ByteBuffer, refIdx: ByteBuffer, uStart: ByteBuffer,
                             uEnd: ByteBuffer, flags: ByteBuffer, fixedLen: Int, n: Int, five: ByteBuffer, three: ByteBuffer,
                             metrics160: ByteBuffer, nAligned: ByteBuffer): Unit
    @native def matchBatchDevice(ctx: Long, seq4: Long, seqOff: Long, len: Long, refIdx: Long, uStart: Long, uEnd: Long, flags: Long,
                                 fixedLen: Int, n: Int, five: Long, three: Long, metrics160: ByteBuffer, nAligned: ByteBuffer): Unit
    @native def matchBatch16Device(ctx: Long, seq4: Long, seqOff: Long, len: Long, refIdx: Long, uStart: Long, uEnd: Long, flags: Long,
                                   fixedLen: Int, n: Int, five: Long, three: Long, metrics160: ByteBuffer, nAligned: ByteBuffer): Unit
    @native def result16Unpack(in: ByteBuffer, n: Long, out: ByteBuffer): Unit
    // result helpers, tag strings (PMa:78-88), metrics (IdentifyPrimersMetric.scala:34-89)
    @native def resultScore(results: ByteBuffer, i: Int): Double
    @native def resultSecond(results: ByteBuffer, i: Int): Double
    @native def metricBin(templateType: Int, isFr: Int, r1: Int, r2: Int): Int
    @native def summaryMetrics(metrics160: ByteBuffer, out17: ByteBuffer): Unit
    @native def formatInfo(panel: Long, results: ByteBuffer, i: Int, readFlags: Int): String
    @native def headerComments(): String
    // downstream classification (scripts/post_process_bam.py:56-78)
    @native def classifyPrimerPair(panel: Long, f5: Int, r5: Int, minInsert: Int, maxInsert: Int): Int
    @native def classifyTemplates(panel: Long, flags: ByteBuffer, five: ByteBuffer, n: Int, minInsert: Int, maxInsert: Int,
                                  clsPerRead: ByteBuffer, counts6: ByteBuffer): Unit
    @native def classifyBatchDevice(ctx: Long, flags: Long, five: Long, compact: Int, n: Int, minInsert: Int, maxInsert: Int,
                                    clsPerRead: Long, counts6: ByteBuffer): Unit
    // BAM records <-> batches
    @native def bamScan(bam: ByteBuffer, nBytes: Long, out2: ByteBuffer): Unit
    @native def bamToBatch(bam: ByteBuffer, nBytes: Long, seq4: ByteBuffer, seqOff: ByteBuffer, len: ByteBuffer, refIdx: ByteBuffer,
                           uStart: ByteBuffer, uEnd: ByteBuffer, flags: ByteBuffer, recIndex: ByteBuffer): Unit
    @native def bamToBatchWindow(panel: Long, bam: ByteBuffer, nBytes: Long, seq4: ByteBuffer, seqOff: ByteBuffer, len: ByteBuffer,
                                 refIdx: ByteBuffer, uStart: ByteBuffer, uEnd: ByteBuffer, flags: ByteBuffer, recIndex: ByteBuffer): Unit
    @native def bamAnnotate(panel: Long, bam: ByteBuffer, nBytes: Long, five: ByteBuffer, three: ByteBuffer, n: Int, out: ByteBuffer, outCap: Long): Long
    @native def packBases(ascii: ByteBuffer, n: Int, seq4: ByteBuffer): Unit
    @native def unpackBases(seq4: ByteBuffer, n: Int, ascii: ByteBuffer): Unit
    // the aligner seam on an existing context (AsyncCudaAligner has its own handle)
    @native def alignBatch(ctx: Long, mode: Int, queries: ByteBuffer, qOff: ByteBuffer, targets: ByteBuffer, tOff: ByteBuffer, n: Int,
                           score: ByteBuffer, targetStart: ByteBuffer, targetEnd: ByteBuffer, queryStart: ByteBuffer, queryEnd: ByteBuffer): Unit
  }
  val ResultBytes = 16 // sizeof(idp_result16): what crosses PCIe
  private def direct(n: Int) = ByteBuffer.allocateDirect(math.max(n, 8)).order(ByteOrder.LITTLE_ENDIAN)
}

/** One per worker thread (an idp_ctx is not thread-safe); the panel handle is shared (IdentifyPrimers.scala:223-234). */
private[identifyprimers] class GpuPrimerMatcher(panel: Long, device: Int, primers: IndexedSeq[Primer], threePrime: Int) extends AutoCloseable {
  import GpuPrimerMatcher._
  private val ctx = Native.ctxCreate(panel, device)

  /** Replaces toPrimeMatchFuture + matchThreePrime for a whole batch; returns per read (fivePrime, threePrime). */
  def matchBatch(templates: Seq[Template]): IndexedSeq[(Option[PrimerMatch], Option[PrimerMatch])] = {
    val recs: IndexedSeq[SamRecord] = templates.flatMap(t => t.r1.toSeq ++ t.r2.toSeq).toIndexedSeq
    val n = recs.length
    // SoA packing: BAM's own 4-bit `seq` bytes go in unchanged (htsjdk SAMRecord#getVariableBinaryRepresentation / bytesToCompressedBases)
    val seqOff = direct(8 * n); val len = direct(4 * n); val refIdx = direct(4 * n); val uS = direct(4 * n); val uE = direct(4 * n); val flags = direct(2 * n)
    // without --three-prime only the leading bases (sequencing order) can matter: ship the first `window` stored bases, or the
    // last ones of a mapped reverse-strand read (PrimerMatcher.scala:39-45); same results, a quarter of the bytes over PCIe
    val window = { val w = Native.panelWindow(panel); if (w > 0) w else Int.MaxValue }
    def shipped(r: SamRecord): Array[Byte] =
      if (r.length <= window) r.bases else if (r.mapped && r.negativeStrand) r.bases.takeRight(window) else r.bases.take(window)
    var off = 0L
    recs.zipWithIndex.foreach { case (r, i) =>
      val kept = math.min(r.length, window)
      seqOff.putLong(8 * i, off); len.putInt(4 * i, kept); off += (kept + 1) / 2
      refIdx.putInt(4 * i, if (r.mapped) r.refIndex else -1)
      uS.putInt(4 * i, if (r.mapped) r.unclippedStart else 0); uE.putInt(4 * i, if (r.mapped) r.unclippedEnd else 0)
      var f = r.flags & (0x1 | 0x4 | 0x10 | 0x40 | 0x80)
      if (r.paired && r.firstOfPair && r.isFrPair) f |= 0x1000
      flags.putShort(2 * i, f.toShort)
    }
    // IDP_F_MATE_NEXT on every R1 that has an R2 in its template
    var i = 0
    templates.foreach { t => if (t.r1.isDefined && t.r2.isDefined) flags.putShort(2 * i, (flags.getShort(2 * i) | 0x2000).toShort); i += t.r1.size + t.r2.size }
    val seq4 = direct(off.toInt)
    recs.zipWithIndex.foreach { case (r, j) => seq4.position(seqOff.getLong(8 * j).toInt); seq4.put(htsjdk.samtools.SAMUtils.bytesToCompressedBases(shipped(r))) }
    val five = direct(ResultBytes * n); val three = if (threePrime >= 0) direct(ResultBytes * n) else null
    val metrics = direct(8 * 160); val nAligned = direct(8)
    Native.matchBatch16(ctx, seq4, seqOff, len, refIdx, uS, uE, flags, 0, n, five, three, metrics, nAligned)
    (0 until n).map { j => (decode(five, j), if (three == null) None else decode(three, j)) }
  }

  /** idp_result16 -> the case classes of PrimerMatch.scala:96-117; a non-zero status is the require() the JVM would have thrown. */
  private def decode(buf: ByteBuffer, j: Int): Option[PrimerMatch] = {
    val b = ResultBytes * j
    val head = buf.getInt(b)
    val primer = (head & 0xFFFFFF) - 1; val kind = (head >>> 24) & 3; val multi = ((head >>> 26) & 1) != 0; val status = (head >>> 27) & 7
    require(status == 0, s"primer matching would have failed a requirement (status $status)")
    val a = buf.getShort(b + 4).toInt; val bb = buf.getShort(b + 6).toInt
    kind match {
      case 0 => None
      case 1 => Some(LocationBasedPrimerMatch(primers(primer), a))
      case 2 => Some(UngappedAlignmentPrimerMatch(primers(primer), a, if (bb == 32767) Int.MaxValue else bb))  // IDP_NEXT16_NA -> "na"
      case 3 =>
        val qLen = buf.getShort(b + 8) & 0xFFFF; val qLenNext = buf.getShort(b + 10) & 0xFFFF
        val score  = if (multi) a / qLen.toDouble else a.toDouble                 // PrimerMatcher.scala:321 vs :315
        val second = if (multi) bb / qLenNext.toDouble else 0d                    // :322
        Some(GappedAlignmentPrimerMatch(primers(primer), score, second, buf.getShort(b + 12) & 0xFFFF, buf.getShort(b + 14) & 0xFFFF))
    }
  }
  override def close(): Unit = Native.ctxDestroy(ctx)
}
