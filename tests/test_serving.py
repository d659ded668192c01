"""TF-Serving-compatible front: wire contract of RecForYouProcess.callNeuralCFTFServing."""
import json
import threading
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
