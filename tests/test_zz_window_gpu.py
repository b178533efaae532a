"""GPU check of the windowed ingest (idp_bam_to_batch_window): the batch cut to the bases the 5' chain can reach gives the same
results as the whole reads, through the CUDA path.  (The CPU proof against the oracle is in test_bam_ingest.py; this file sorts
last on purpose: it is the newest path.)"""
import numpy as np
import pytest

import helpers as H
import fg_idprimer_b200 as idp
from test_bam_ingest import _window_panel, _window_reads, recs_to_bam

pytestmark = pytest.mark.gpu
FIELDS = ["primer", "type", "status", "multi", "mm", "next_mm", "raw_score", "raw_next", "q_len", "q_len_next", "t_start", "t_end"]


@pytest.mark.parametrize("opts", [dict(ungapped_kmer_length=6, gapped_kmer_length=10), dict(slop=2, max_mismatch_rate=0.1)])
def test_windowed_batch_on_gpu(opts):
    rng = np.random.default_rng(23)
    primers = _window_panel(rng, 12, 18, 27)
    recs = _window_reads(rng, primers, 3000, [8, 20, 33, 34, 35, 39, 40, 41, 75, 150])
    for i in range(0, len(recs) - 1, 3):
        recs[i].paired = recs[i + 1].paired = True
        recs[i].first, recs[i + 1].second, recs[i].mate_next = True, True, True
    bam = recs_to_bam(recs)
    gp = [idp.Primer(p.pair_id, p.primer_id, p.sequence, p.positive_strand, p.ref_name, p.start, p.end) for p in primers]
    with idp.IdentifyPrimers(gp, ref_names=H.DEFAULT_DICT, **opts) as tool:
        whole, _ = idp.ReadBatch.from_bam(bam)
        cut, _ = idp.ReadBatch.from_bam(bam, window_for=tool)
        assert cut.seq4.nbytes < whole.seq4.nbytes and int(cut.len.max()) == tool.window()
        want5, _ = tool.process_batch(whole)
        want5 = want5.copy()
        want_metrics, want_align = tool.metric_counter.copy(), tool.num_alignments
        got5, _ = tool.process_batch(cut)
        for f in FIELDS:
            assert np.array_equal(got5[f], want5[f]), f
        # process_batch accumulates like the tool's counter does: the second batch doubles both
        assert np.array_equal(tool.metric_counter, 2 * want_metrics) and tool.num_alignments == 2 * want_align
        # the same two batches through the 16-byte records (idp_match_batch16, what the plugin ships by default): after
        # idp_result16_unpack the rows are the very bytes of the 32-byte records
        for batch in (whole, cut):
            c5, _ = tool.process_batch(batch, compact=True)
            assert c5.tobytes() == want5.tobytes()
        assert np.array_equal(tool.metric_counter, 4 * want_metrics) and tool.num_alignments == 4 * want_align
