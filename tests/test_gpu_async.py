"""bsw_extend_async / bsw_wait: many small calls from many threads are coalesced into shared batches;
every call gets exactly the results of a synchronous call, errors stay with the call that caused them."""
import threading

import numpy as np
import pytest

from conftest import load_golden, results_matrix

pytestmark = pytest.mark.gpu


def test_async_calls_coalesce_and_match_sync(lib):
    cfg = lib.gen_named_config("small")
    pairs, ref, qer = lib.gen_pairs(cfg, 0, 40 * 512)
    want = pairs.copy()
    with lib.Engine() as eng:
        eng.extend(want, ref, qer, 100)
        # 1) submit everything, then wait: the worker finds many calls queued at once
        got = pairs.copy()
        views = [got[k * 512:(k + 1) * 512] for k in range(40)]
        tickets = [eng.extend_async(v, ref, qer, 100) for v in views]
        cells = sum(eng.wait(t) for t in tickets)
        assert np.array_equal(results_matrix(got), results_matrix(want))
        calls, batches = eng.async_stats()
        assert calls == 40 and batches < calls, (calls, batches)
        assert cells > 0
        # 2) eight threads, blocking submit + wait per 512-pair call (the driver's -t 8 -b 512 shape)
        got2 = pairs.copy()
        errs = []

        def worker(tid):
            try:
                for k in range(tid, 40, 8):
                    v = got2[k * 512:(k + 1) * 512]
                    eng.wait(eng.extend_async(v, ref, qer, 100))
            except Exception as e:                                # pragma: no cover
                errs.append(e)
        ts = [threading.Thread(target=worker, args=(t,)) for t in range(8)]
        [t.start() for t in ts]
        [t.join() for t in ts]
        assert not errs
        assert np.array_equal(results_matrix(got2), results_matrix(want))
        # the engine itself stays usable for synchronous calls while the queue exists
        again = pairs.copy()
        eng.extend(again, ref, qer, 100)
        assert np.array_equal(results_matrix(again), results_matrix(want))


def test_async_different_buffers_bands_and_errors(lib):
    """Calls over different sequence buffers and different w in one queue; a bad call fails alone."""
    pa, ra, qa, wa, params, expa, _ = load_golden("short8")
    pb, rb, qb, wb, _, expb, _ = load_golden("small_151bp")
    with lib.Engine(**params) as eng:
        a1, a2, b1 = pa.copy(), pa.copy(), pb.copy()
        bad = pa[:300].copy()
        bad["len1"][7] = 0
        t = [eng.extend_async(a1, ra, qa, wa), eng.extend_async(b1, rb, qb, wb), eng.extend_async(bad, ra, qa, wa),
             eng.extend_async(a2, ra, qa, wa), eng.extend_async(pa[:0].copy(), ra, qa, wa)]
        eng.wait(t[0]); eng.wait(t[1])
        with pytest.raises(lib.BswError) as ei:
            eng.wait(t[2])
        assert ei.value.code == -2
        eng.wait(t[3]); eng.wait(t[4])
        with pytest.raises(lib.BswError) as ei:
            eng.wait(t[3])                                        # a ticket can be waited for once
        assert ei.value.code == -5
        assert np.array_equal(results_matrix(a1), expa) and np.array_equal(results_matrix(a2), expa)
        assert np.array_equal(results_matrix(b1), expb)
