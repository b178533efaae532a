/*
 * Reference-side binding (NOT compiled in this image: no JVM).  Lives in package com.fulcrumgenomics.identifyprimers so that it can
 * see the private[identifyprimers] types, and replaces the body of IdentifyPrimers.processBatch (IdentifyPrimers.scala:381-436):
 * the per-read Future plumbing becomes ONE native call per 5,000-template batch.
 */
package com.fulcrumgenomics.identifyprimers

import java.nio.{ByteBuffer, ByteOrder}

import com.fulcrumgenomics.bam.Template
import com.fulcrumgenomics.bam.api.SamRecord

private[identifyprimers] object GpuPrimerMatcher {
  object Native {
    System.loadLibrary("idprimer_jni")
    @native def panelCreate(primers: ByteBuffer, n: Int, params: ByteBuffer): Long
    @native def panelDestroy(h: Long): Unit
    @native def panelWindow(panel: Long): Int  // idp_panel_window: bases the 5' chain can reach, -1 with --three-prime
    @native def ctxCreate(panel: Long, device: Int): Long
    @native def ctxDestroy(h: Long): Unit
    @native def matchBatch(ctx: Long, seq4: ByteBuffer, seqOff: ByteBuffer, len: ByteBuffer, refIdx: ByteBuffer, uStart: ByteBuffer,
                           uEnd: ByteBuffer, flags: ByteBuffer, fixedLen: Int, n: Int, five: ByteBuffer, three: ByteBuffer,
                           metrics160: ByteBuffer, nAligned: ByteBuffer): Unit
  }
  val ResultBytes = 32 // sizeof(idp_result)
  private def direct(n: Int) = ByteBuffer.allocateDirect(n).order(ByteOrder.LITTLE_ENDIAN)
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
    Native.matchBatch(ctx, seq4, seqOff, len, refIdx, uS, uE, flags, 0, n, five, three, metrics, nAligned)
    (0 until n).map { j => (decode(five, j), if (three == null) None else decode(three, j)) }
  }

  /** idp_result -> the case classes of PrimerMatch.scala:96-117; a non-zero status is the require() the JVM would have thrown. */
  private def decode(buf: ByteBuffer, j: Int): Option[PrimerMatch] = {
    val b = ResultBytes * j
    val status = buf.get(b + 5)
    require(status == 0, s"primer matching would have failed a requirement (status $status)")
    buf.get(b + 4) match {
      case 0 => None
      case 1 => Some(LocationBasedPrimerMatch(primers(buf.getInt(b)), buf.getInt(b + 8)))
      case 2 => Some(UngappedAlignmentPrimerMatch(primers(buf.getInt(b)), buf.getInt(b + 8), buf.getInt(b + 12)))
      case 3 =>
        val multi = buf.get(b + 6) != 0
        val raw = buf.getInt(b + 16); val rawNext = buf.getInt(b + 20); val qLen = buf.getShort(b + 24); val qLenNext = buf.getShort(b + 26)
        val score  = if (multi) raw / qLen.toDouble else raw.toDouble            // PrimerMatcher.scala:321 vs :315
        val second = if (multi) rawNext / qLenNext.toDouble else 0d               // :322
        Some(GappedAlignmentPrimerMatch(primers(buf.getInt(b)), score, second, buf.getShort(b + 28), buf.getShort(b + 30)))
    }
  }
  override def close(): Unit = Native.ctxDestroy(ctx)
}
