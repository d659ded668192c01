"""TF-Serving-compatible front: wire contract of RecForYouProcess.callNeuralCFTFServing."""
import json
import threading
import time
import urllib.error
import urllib.request

import numpy as np
import pytest

from conftest import load_golden_weights
from sparrowrecsys_b200 import serving
from sparrowrecsys_b200.spec import default_spec


def _post(port, path, payload):
    req = urllib.request.Request("http://127.0.0.1:%d%s" % (port, path), data=json.dumps(payload).encode(),
                                 headers={"Content-Type": "application/json"})
    try:
        with urllib.request.urlopen(req, timeout=30) as r:
            return r.status, json.loads(r.read())
    except urllib.error.HTTPError as e:
        return e.code, json.loads(e.read())


def _run(models):
    srv = serving.serve(models, port=0)
    t = threading.Thread(target=srv.serve_forever, daemon=True)
    t.start()
    return srv, srv.server_address[1]


def test_row_format_to_columns():
    spec = default_spec("din")
    f = serving.instances_to_features(spec, [{"userId": 3, "movieId": 7, "userGenre1": "Drama",
                                              "movieAvgRating": 3.5}, {"userId": 4, "movieId": 8}])
    assert f["userId"].tolist() == [3, 4] and f["movieId"].dtype == np.int64
    assert f["userGenre1"].tolist() == ["Drama", ""]          # missing string -> ""
    assert f["userRatedMovie3"].tolist() == [0, 0]            # missing int -> 0
    assert f["movieAvgRating"].dtype == np.float32 and f["movieAvgRating"][1] == 0.0
    with pytest.raises(ValueError):
        serving.instances_to_features(spec, [])


def test_wire_contract_with_stand_in_scorer():
    """Host logic only (no GPU): the scorer is the oracle on the shipped NeuralCF weights."""
    from oracle import ctr_oracle as O
    spec = default_spec("neuralcf")
    W = load_golden_weights("neuralcf_002")
    srv, port = _run({"recmodel": (spec, lambda f: O.forward(spec, W, f)[0])})
    try:
        # the request online/util/HttpClient.java:110-147 builds
        code, body = _post(port, "/v1/models/recmodel:predict",
                           {"instances": [{"userId": 10351, "movieId": 52}, {"userId": 10351, "movieId": 53}]})
        assert code == 200
        preds = body["predictions"]
        assert len(preds) == 2 and len(preds[0]) == 1            # read as predictions[i][0] in Java
        assert abs(preds[0][0] - 0.68536943) < 1e-6 and abs(preds[1][0] - 0.17321654) < 1e-6
        code, body = _post(port, "/v1/models/recmodel:predict", {"instances": [{"userId": 99999, "movieId": 1}]})
        assert code == 400 and "error" in body                   # out-of-range id
        code, body = _post(port, "/v1/models/other:predict", {"instances": [{"userId": 1, "movieId": 1}]})
        assert code == 404
        code, body = _post(port, "/v1/models/recmodel:predict", {"nope": 1})
        assert code == 400
        # TF-Serving's columnar format is answered with "outputs"
        code, body = _post(port, "/v1/models/recmodel:predict",
                           {"inputs": {"userId": [10351, 10351], "movieId": [[52], [53]]}})
        assert code == 200 and abs(body["outputs"][0][0] - 0.68536943) < 1e-6 \
            and abs(body["outputs"][1][0] - 0.17321654) < 1e-6
        code, body = _post(port, "/v1/models/recmodel:predict", {"inputs": {"userId": [1, 2], "movieId": [3]}})
        assert code == 400
    finally:
        srv.shutdown()


@pytest.mark.gpu
def test_http_front_on_the_cuda_model():
    from tfrecmodel import neuralcf
    model = neuralcf.load(weights=load_golden_weights("neuralcf_002"))
    srv, port = _run({"recmodel": (model.spec, model.predict)})
    try:
        inst = [{"userId": 10351, "movieId": m} for m in (52, 53)] + \
               [{"userId": 14887, "movieId": 1}]
        code, body = _post(port, "/v1/models/recmodel:predict", {"instances": inst})
        assert code == 200
        got = [p[0] for p in body["predictions"]]
        np.testing.assert_allclose(got, [0.68536943, 0.17321654, 0.8525178], atol=1e-6)
        many = [{"userId": 14887, "movieId": 1 + (i % 900)} for i in range(800)]   # CANDIDATE_SIZE
        code, body = _post(port, "/v1/models/recmodel:predict", {"instances": many})
        assert code == 200 and len(body["predictions"]) == 800
    finally:
        srv.shutdown()
        model.close()


# ---- cross-request micro-batching (host logic, stand-in scorer) ---------------------------
def _feats(ids):
    ids = np.asarray(ids, np.int64)
    return {"movieId": ids, "userId": ids * 10}


def test_micro_batcher_merges_requests_that_wait_behind_a_call():
    gate, first_in = threading.Event(), threading.Event()
    seen = []

    def scorer(f):
        seen.append(f["movieId"].tolist())
        if len(seen) == 1:
            first_in.set()
            gate.wait(10)
        return (f["movieId"] + f["userId"] * 0.001).astype(np.float32).reshape(-1, 1)

    mb = serving.MicroBatcher(scorer, max_rows=1000)
    out = {}
    worker = lambda name, ids: out.__setitem__(name, mb.submit(_feats(ids)))
    t0 = threading.Thread(target=worker, args=("a", [1, 2]))
    t0.start()
    assert first_in.wait(10)                        # call 1 is running with request a alone
    rest = [threading.Thread(target=worker, args=(n, ids))
            for n, ids in (("b", [3]), ("c", [4, 5, 6]), ("d", [7]))]
    for t in rest:
        t.start()
    deadline = time.monotonic() + 10
    while mb._q.qsize() < 3 and time.monotonic() < deadline:
        time.sleep(0.005)
    gate.set()
    for t in [t0] + rest:
        t.join(10)
    mb.close()
    assert seen[0] == [1, 2] and len(seen) == 2 and sorted(seen[1]) == [3, 4, 5, 6, 7]
    assert mb.calls == 2 and mb.requests == 4
    for name, ids in (("a", [1, 2]), ("b", [3]), ("c", [4, 5, 6]), ("d", [7])):
        exp = np.array(ids, np.float32) * np.float32(1.01)
        assert out[name].shape == (len(ids), 1)
        np.testing.assert_allclose(out[name][:, 0], exp, rtol=1e-6)


def test_micro_batcher_respects_max_rows_and_isolates_errors():
    gate, first_in = threading.Event(), threading.Event()
    sizes = []

    def scorer(f):
        sizes.append(len(f["movieId"]))
        if len(sizes) == 1:
            first_in.set()
            gate.wait(10)
        if (f["movieId"] < 0).any():
            raise ValueError("movie id out of range")
        return f["movieId"].astype(np.float32).reshape(-1, 1)

    mb = serving.MicroBatcher(scorer, max_rows=4)
    res = {}

    def worker(name, ids):
        try:
            res[name] = mb.submit(_feats(ids))
        except ValueError as e:
            res[name] = e

    t0 = threading.Thread(target=worker, args=("w", [9]))
    t0.start()
    assert first_in.wait(10)
    ts = []
    for name, ids in (("x", [1, 2]), ("bad", [-1]), ("y", [3]), ("z", [4, 5, 6])):
        t = threading.Thread(target=worker, args=(name, ids))
        t.start()
        ts.append(t)
        while mb._q.qsize() < len(ts):              # keep the arrival order deterministic
            time.sleep(0.002)
    gate.set()
    for t in [t0] + ts:
        t.join(10)
    mb.close()
    assert isinstance(res["bad"], ValueError)
    assert res["x"][:, 0].tolist() == [1, 2] and res["y"][:, 0].tolist() == [3]
    assert res["z"][:, 0].tolist() == [4, 5, 6] and res["w"][:, 0].tolist() == [9]
    # call 1: w; call 2: x+bad+y merged (4 rows) fails -> re-run one by one; then z alone
    assert sizes == [1, 4, 2, 1, 1, 3]
    with pytest.raises(RuntimeError):
        mb.submit(_feats([1]))


def test_concurrent_http_requests_share_library_calls():
    """Many Jetty-style blocking POSTs at once: fewer scorer calls than requests, every
    response in its own request's order."""
    from oracle import ctr_oracle as O
    spec = default_spec("neuralcf")
    W = load_golden_weights("neuralcf_002")
    lock = threading.Lock()

    def scorer(f):
        with lock:
            time.sleep(0.02)          # a call long enough for requests to queue
            return O.forward(spec, W, f)[0]

    srv, port = _run({"recmodel": (spec, scorer)})
    try:
        results = {}

        def client(i):
            inst = [{"userId": 10351, "movieId": 52 + ((i + j) % 2)} for j in range(3)]
            results[i] = _post(port, "/v1/models/recmodel:predict", {"instances": inst})

        threads = [threading.Thread(target=client, args=(i,)) for i in range(16)]
        for t in threads:
            t.start()
        for t in threads:
            t.join(30)
        golden = {52: 0.68536943, 53: 0.17321654}
        for i in range(16):
            code, body = results[i]
            assert code == 200
            exp = [golden[52 + ((i + j) % 2)] for j in range(3)]
            np.testing.assert_allclose([p[0] for p in body["predictions"]], exp, atol=1e-6)
        mb = srv.batchers["recmodel"]
        assert mb.requests == 16 and mb.calls < 16
    finally:
        srv.shutdown()
        srv.server_close()
    assert not mb._thread.is_alive()                          # dispatcher stopped with the server


def test_http_front_answers_every_failure_with_a_json_error():
    """TF-Serving answers every failure with {"error": ...} (HttpClient.java:32-40 reads the body):
    a library failure is a 500, a non-object body a 400 - never a dropped connection."""
    import json
    import threading
    import urllib.error
    import urllib.request
    from sparrowrecsys_b200 import serving
    from sparrowrecsys_b200.spec import default_spec

    def broken(feats):
        raise RuntimeError("CUDA went away")

    srv = serving.serve({"recmodel": (default_spec("neuralcf"), broken)}, port=0, micro_batch=True)
    port = srv.server_address[1]
    th = threading.Thread(target=srv.serve_forever, daemon=True)
    th.start()
    try:
        def post(body):
            req = urllib.request.Request("http://127.0.0.1:%d/v1/models/recmodel:predict" % port, data=body,
                                         headers={"Content-Type": "application/json"})
            try:
                with urllib.request.urlopen(req, timeout=10) as r:
                    return r.status, json.loads(r.read())
            except urllib.error.HTTPError as e:
                return e.code, json.loads(e.read())
        code, body = post(json.dumps({"instances": [{"userId": 1, "movieId": 2}]}).encode())
        assert code == 500 and "CUDA went away" in body["error"]
        code, body = post(b"[1, 2, 3]")
        assert code == 400 and "error" in body
    finally:
        srv.shutdown()
        srv.server_close()


def test_micro_batcher_submit_after_close_raises_instead_of_hanging():
    import pytest as _pt
    from sparrowrecsys_b200.serving import MicroBatcher
    mb = MicroBatcher(lambda f: np.zeros((len(f["movieId"]), 1), np.float32))
    assert mb.submit({"movieId": np.arange(3)}).shape == (3, 1)
    mb.close()
    with _pt.raises(RuntimeError):
        mb.submit({"movieId": np.arange(3)})
