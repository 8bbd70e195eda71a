"""CPU: the server's GPU-thread loop (mlx_vlm_b200/server.py; reference server/generation.py:1730-1918)
driven with the deterministic engine stand-in of test_batch_host.py: concurrent callers get exactly the
tokens their request produces alone, the admission cap holds, a cancelled request frees its row without
disturbing the others, a failing request is reported to its caller only, and stop drains cleanly."""
import threading
import types

import numpy as np
import pytest
import torch

from test_batch_host import _alone, _fake_model


def _server(max_num_seqs=16, lockstep=True, **kw):
    from mlx_vlm_b200.server import ResponseGenerator
    model, proc = _fake_model(lockstep=lockstep)
    inner = model.get_input_embeddings

    def get_input_embeddings(ids, pixel_values=None, mask=None, **k):
        if k.get("boom"):
            raise ValueError("bad request")
        e = inner(np.asarray(ids).reshape(1, -1), pixel_values, mask)
        e.to_dict = lambda: {"inputs_embeds": e.inputs_embeds, "position_ids": e.position_ids,
                             "rope_deltas": e.rope_deltas}
        return e
    model.get_input_embeddings = get_input_embeddings
    return ResponseGenerator(model, proc, max_num_seqs=max_num_seqs, decode_slice=4, **kw), model


PROMPTS = [[5, 6, 7, 8, 9], [1, 2], [3, 3, 3], [11] * 9, [4, 4], [9, 8, 7, 6]]
IMAGES = [True, False, True, False, False, True]
MAXES = [7, 12, 1, 9, 5, 10]


@pytest.mark.parametrize("lockstep", [True, False])
def test_concurrent_requests_get_their_own_tokens(lockstep):
    from mlx_vlm_b200.server import GenerationArguments
    srv, _ = _server(lockstep=lockstep)
    out = {}

    def client(i):
        raw = {"input_ids": np.asarray([PROMPTS[i]])}
        if IMAGES[i]:
            raw["pixel_values"] = 1
        evs = list(srv.generate(raw, GenerationArguments(max_tokens=MAXES[i]), timeout=20))
        out[i] = ([e.token for e in evs], evs[-1].finish_reason)

    th = [threading.Thread(target=client, args=(i,)) for i in range(len(PROMPTS))]
    for t in th:
        t.start()
    for t in th:
        t.join(30)
    srv.stop_and_join()
    assert not srv._thread.is_alive()
    for i in range(len(PROMPTS)):
        assert out[i] == _alone(PROMPTS[i], MAXES[i], IMAGES[i]), i


def test_admission_cap_and_backpressure():
    from mlx_vlm_b200.server import GenerationArguments
    srv, _ = _server(max_num_seqs=2)
    reqs = [srv.submit({"input_ids": np.asarray([p])}, GenerationArguments(max_tokens=m))
            for p, m in zip(PROMPTS, MAXES)]
    got = []
    for r, p, m in zip(reqs, PROMPTS, MAXES):
        toks = []
        while True:
            item = r.rqueue.get(timeout=20)
            if item is None:
                break
            if hasattr(item, "token"):
                toks.append(item.token)
        got.append(toks)
        assert toks == _alone(p, m)[0]
    srv.stop_and_join()
    assert srv.peak_active <= 2 and srv.steps > 0


def test_cancel_and_error_isolation():
    from mlx_vlm_b200.server import GenerationArguments, GenerationContext
    srv, _ = _server()
    long_req = srv.submit({"input_ids": np.asarray([[7, 7, 7]])}, GenerationArguments(max_tokens=100000))
    ctx = long_req.rqueue.get(timeout=20)
    assert isinstance(ctx, GenerationContext) and ctx.prompt_tokens == 3
    first = long_req.rqueue.get(timeout=20)
    assert hasattr(first, "token")
    bad = srv.submit({"input_ids": np.asarray([[1, 2, 3]]), "boom": True}, GenerationArguments(max_tokens=4))
    err = bad.rqueue.get(timeout=20)
    assert isinstance(err, ValueError)
    ok = list(srv.generate({"input_ids": np.asarray([[1, 2]])}, GenerationArguments(max_tokens=12), timeout=20))
    assert [e.token for e in ok] == _alone([1, 2], 12)[0]        # unaffected by the neighbour and the failure
    srv.cancel(long_req)
    while True:                                                   # the cancelled stream ends with None
        item = long_req.rqueue.get(timeout=20)
        if item is None:
            break
    with pytest.raises(ValueError):
        list(srv.generate({"input_ids": np.asarray([[1]]), "boom": True}, timeout=20))
    srv.stop_and_join()
    assert not srv._thread.is_alive()
    with pytest.raises(RuntimeError):
        srv.submit({"input_ids": np.asarray([[1]])})
