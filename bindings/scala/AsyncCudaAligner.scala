/*
 * Reference-side binding (NOT compiled in this image: no JVM).  A batched, GPU-backed AsyncAligner for
 * src/main/scala/com/fulcrumgenomics/alignment/AsyncAligner.scala: the trait it implements is AsyncAligner (AA:37-54), the
 * precedent for an out-of-JVM backend is AsyncKswAligner (AA:112-196) -- same constructor surface (scores + mode), Closeable,
 * alignments complete in the order they were added, a failure surfaces as a failed Future / IllegalStateException.
 *
 * IdentifyPrimers creates one aligner per processBatch call (IP:364-372, 385-386), submits every task of the batch and only
 * then awaits the futures (IP:441-452, 404-411).  align() therefore only queues; a batch goes to the GPU in ONE
 * idp_align_batch call (include/idprimer.h) when it is full, when nothing new has been queued for `lingerMillis`, or on
 * flush()/close().  newAligner (IP:364-372) gains one case:   if (gpu) new AsyncCudaAligner(matchScore, mismatchScore, gapOpen,
 * gapExtend, mode) ...
 */
package com.fulcrumgenomics.alignment

import java.io.Closeable
import java.nio.{ByteBuffer, ByteOrder}
import java.util.concurrent.atomic.AtomicLong
import java.util.concurrent.{Executors, ScheduledExecutorService, ThreadFactory, TimeUnit}

import com.fulcrumgenomics.alignment.Mode.{Global, Glocal, Local}

import scala.collection.mutable.ArrayBuffer
import scala.concurrent.{ExecutionContext, Future, Promise}

object AsyncCudaAligner {
  /** The native half: bindings/jni/idprimer_jni.c. */
  object Native {
    System.loadLibrary("idprimer_jni")
    /** A context whose panel is a two-primer placeholder: idp_align_batch only uses the scoring parameters. */
    @native def alignerCreate(matchScore: Int, mismatchScore: Int, gapOpen: Int, gapExtend: Int, device: Int): Long
    @native def alignerDestroy(handle: Long): Unit
    /** idp_align_batch: queries / targets are ASCII, concatenated, with n + 1 int64 offsets; every buffer is direct. */
    @native def alignBatch(handle: Long, mode: Int, queries: ByteBuffer, qOff: ByteBuffer, targets: ByteBuffer, tOff: ByteBuffer, n: Int,
                           score: ByteBuffer, targetStart: ByteBuffer, targetEnd: ByteBuffer, queryStart: ByteBuffer, queryEnd: ByteBuffer): Unit
  }

  private final case class Task(query: Array[Byte], target: Array[Byte], promise: Promise[Alignment])

  private val flusherThreads: ThreadFactory = new ThreadFactory {
    override def newThread(r: Runnable): Thread = { val t = new Thread(r, "AsyncCudaAligner-flusher"); t.setDaemon(true); t }
  }
}

/** An asynchronous aligner that runs batches of alignments on a B200 through libidprimer_b200.so.
  *
  * IMPORTANT: like [[AsyncKswAligner]], the alignments are completed in the order added.
  *
  * The [[Alignment]]s returned carry the score, queryStart / queryEnd and targetStart / targetEnd of the optimal alignment
  * under fgbio's [[Aligner]] rules.  The kernels do not materialise a traceback, so the `cigar` is a PLACEHOLDER with the
  * right lengths on query and target (matches, then one insertion or deletion for the difference): everything the primer
  * matchers read (`score`, `targetStart`, `targetEnd`: PrimerMatcher.scala:315-324) is exact, the operator sequence is not.
  *
  * @param matchScore the match score (should be greater than or equal to zero).
  * @param mismatchScore the mismatch score (should be less than or equal to zero).
  * @param gapOpen the gap opening penalty, should generally be negative or zero
  * @param gapExtend the gap extension penalty, should generally be negative or zero
  * @param mode alignment mode to use when generating alignments
  * @param device the CUDA device ordinal
  * @param maxBatch alignments per native call
  * @param lingerMillis a partial batch is sent when nothing has been added for this long
  */
class AsyncCudaAligner(matchScore: Int,
                       mismatchScore: Int,
                       gapOpen: Int,
                       gapExtend: Int,
                       mode: Mode = Mode.Glocal,
                       device: Int = 0,
                       maxBatch: Int = 65536,
                       lingerMillis: Long = 1) extends AsyncAligner with Closeable {
  import AsyncCudaAligner._

  private val handle  = Native.alignerCreate(matchScore, mismatchScore, gapOpen, gapExtend, device)
  private val aligned = new AtomicLong(0)
  private val lock    = new Object
  private var pending = new ArrayBuffer[Task](maxBatch)
  private var closed  = false
  private var lastAdd = System.nanoTime()
  // numbered like the ksw mapping at AA:124-128, which is what idp_align_batch takes (IDP_MODE_*)
  private val modeCode = mode match { case Local => 0; case Glocal => 1; case Global => 3 }

  private val flusher: ScheduledExecutorService = Executors.newSingleThreadScheduledExecutor(flusherThreads)
  flusher.scheduleWithFixedDelay(new Runnable {
    override def run(): Unit = {
      val batch = lock.synchronized {
        if (pending.nonEmpty && System.nanoTime() - lastAdd >= lingerMillis * 1000000L) take() else null
      }
      if (batch != null) run(batch)
    }
  }, lingerMillis, lingerMillis, TimeUnit.MILLISECONDS)

  /** Queues one alignment; the Future completes when its batch has run on the GPU. */
  def align(query: Array[Byte], target: Array[Byte])(implicit e: ExecutionContext): Future[Alignment] = {
    val promise = Promise[Alignment]()
    val full = lock.synchronized {
      if (closed) return Future.failed(new IllegalStateException("AsyncCudaAligner is closed but align() was called"))
      pending += Task(query, target, promise)
      lastAdd = System.nanoTime()
      if (pending.length >= maxBatch) take() else null
    }
    if (full != null) run(full)
    promise.future
  }

  /** Sends whatever is queued now (processBatch may call this once after its last align()). */
  def flush(): Unit = { val batch = lock.synchronized(if (pending.nonEmpty) take() else null); if (batch != null) run(batch) }

  /** The number of completed alignments. */
  override def numAligned: Long = aligned.get()

  /** Completes everything queued, then releases the GPU context. */
  override def close(): Unit = {
    flush()
    lock.synchronized { closed = true }
    flusher.shutdownNow()
    runLock.synchronized { Native.alignerDestroy(handle) }
  }

  private def take(): ArrayBuffer[Task] = { val b = pending; pending = new ArrayBuffer[Task](maxBatch); b }

  private val runLock = new Object // an idp_ctx serves one call at a time; batches also complete in submission order this way

  private def direct(bytes: Int): ByteBuffer = ByteBuffer.allocateDirect(math.max(bytes, 8)).order(ByteOrder.LITTLE_ENDIAN)

  private def run(batch: ArrayBuffer[Task]): Unit = runLock.synchronized {
    val n = batch.length
    try {
      val queries = direct(batch.iterator.map(_.query.length).sum)
      val targets = direct(batch.iterator.map(_.target.length).sum)
      val qOff = direct(8 * (n + 1)); val tOff = direct(8 * (n + 1))
      var q = 0L; var t = 0L; var i = 0
      batch.foreach { task =>
        qOff.putLong(8 * i, q); tOff.putLong(8 * i, t)
        queries.put(task.query); targets.put(task.target)
        q += task.query.length; t += task.target.length; i += 1
      }
      qOff.putLong(8 * n, q); tOff.putLong(8 * n, t)
      val score = direct(4 * n); val tStart = direct(4 * n); val tEnd = direct(4 * n); val qStart = direct(4 * n); val qEnd = direct(4 * n)
      Native.alignBatch(handle, modeCode, queries, qOff, targets, tOff, n, score, tStart, tEnd, qStart, qEnd)
      i = 0
      batch.foreach { task =>
        val (qs, qe, ts, te) = (qStart.getInt(4 * i), qEnd.getInt(4 * i), tStart.getInt(4 * i), tEnd.getInt(4 * i))
        task.promise.success(new Alignment(task.query, task.target, qs, ts, placeholderCigar(qe - qs + 1, te - ts + 1), score.getInt(4 * i)))
        aligned.getAndIncrement()
        i += 1
      }
    } catch {
      case ex: Throwable => batch.foreach(_.promise.tryFailure(new IllegalStateException("CUDA aligner error.", ex)))
    }
  }

  /** Lengths on query and target as the real alignment has them; the operators are not the real ones (see the class doc). */
  private def placeholderCigar(lenOnQuery: Int, lenOnTarget: Int): Cigar = {
    val m = math.min(lenOnQuery, lenOnTarget)
    val elems = IndexedSeq.newBuilder[CigarElem]
    if (m > 0) elems += CigarElem(htsjdk.samtools.CigarOperator.M, m)
    if (lenOnQuery > m) elems += CigarElem(htsjdk.samtools.CigarOperator.I, lenOnQuery - m)
    if (lenOnTarget > m) elems += CigarElem(htsjdk.samtools.CigarOperator.D, lenOnTarget - m)
    Cigar(elems.result())
  }
}
