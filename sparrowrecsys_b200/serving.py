"""TF-Serving-compatible REST front for the CTR models.

The only place the reference's online system calls a CTR model is one HTTP POST from
Jetty to TF-Serving (`online/recprocess/RecForYouProcess.java:113-138`):

    POST http://localhost:8501/v1/models/recmodel:predict
    {"instances": [{"userId": 1, "movieId": 52}, ...]}       (row format, <= 800 rows)
    -> {"predictions": [[0.68], [0.17], ...]}                read as predictions[i][0]

This module answers exactly that contract, so the unmodified Java server can point at it:

    python -m sparrowrecsys_b200.serving --model neuralcf --savedmodel <dir> [--port 8501]

Row-format instances are turned into the feature dict of `model.predict` (columns by key;
a key missing from an instance takes the `make_csv_dataset(na_value="0")` defaults: 0 for
numbers, "" for strings), scored in one library call, and returned in request order.
Errors use TF-Serving's shape: HTTP 400 `{"error": "..."}`.
"""
from __future__ import annotations

import argparse
import json
import re
import threading
from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer
from typing import Callable, Dict, List, Mapping

import numpy as np

from .spec import MOVIE_GENRE_KEYS, USER_GENRE_KEYS, ModelSpec

_PATH = re.compile(r"^/v1/models/([^/:]+)(?:/versions/\d+)?:predict$")
_STRING_KEYS = set(MOVIE_GENRE_KEYS) | set(USER_GENRE_KEYS)


def instances_to_features(spec: ModelSpec, instances: List[Mapping[str, object]]) -> Dict[str, np.ndarray]:
    """TF-Serving row format -> dict of columns for `predict`."""
    if not isinstance(instances, list) or not instances:
        raise ValueError("'instances' must be a non-empty list")
    feats: Dict[str, np.ndarray] = {}
    for key in spec.required_keys():
        if key in _STRING_KEYS:
            col = np.array([str(inst.get(key, "")) for inst in instances], dtype=object)
        else:
            col = np.array([float(inst.get(key, 0) or 0) for inst in instances], dtype=np.float64)
            if key.endswith("Id") or key.startswith("userRatedMovie"):
                col = col.astype(np.int64)
            else:
                col = col.astype(np.float32)
        feats[key] = col
    return feats


def make_handler(models: Mapping[str, tuple], lock: threading.Lock):
    """`models`: name -> (spec, predict_fn); predict_fn(features) -> float32 [N,1]."""

    class Handler(BaseHTTPRequestHandler):
        protocol_version = "HTTP/1.1"

        def log_message(self, fmt, *args):       # quiet by default
            pass

        def _send(self, code, payload):
            body = json.dumps(payload).encode()
            self.send_response(code)
            self.send_header("Content-Type", "application/json")
            self.send_header("Content-Length", str(len(body)))
            self.end_headers()
            self.wfile.write(body)

        def do_POST(self):
            m = _PATH.match(self.path)
            if not m or m.group(1) not in models:
                return self._send(404, {"error": "Servable not found for request: %s" % self.path})
            spec, predict_fn = models[m.group(1)]
            try:
                n = int(self.headers.get("Content-Length", "0"))
                req = json.loads(self.rfile.read(n) or b"{}")
                feats = instances_to_features(spec, req.get("instances"))
                with lock:                       # one GPU stream of work at a time per process
                    p = predict_fn(feats)
                self._send(200, {"predictions": [[float(v)] for v in np.asarray(p).reshape(-1)]})
            except (ValueError, KeyError, TypeError, json.JSONDecodeError) as e:
                self._send(400, {"error": str(e)})

        def do_GET(self):
            m = re.match(r"^/v1/models/([^/:]+)$", self.path)
            if m and m.group(1) in models:
                return self._send(200, {"model_version_status": [
                    {"version": "1", "state": "AVAILABLE", "status": {"error_code": "OK", "error_message": ""}}]})
            self._send(404, {"error": "not found"})

    return Handler


def serve(models: Mapping[str, tuple], host: str = "127.0.0.1", port: int = 8501) -> ThreadingHTTPServer:
    """Start the server in the calling thread's process; returns the server object
    (call `.serve_forever()`; `.shutdown()` from another thread stops it)."""
    return ThreadingHTTPServer((host, port), make_handler(models, threading.Lock()))


def main():
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--model", default="neuralcf")
    ap.add_argument("--name", default="recmodel", help="servable name in the URL")
    ap.add_argument("--savedmodel", help="reference SavedModel dir (neuralcf / twotowers)")
    ap.add_argument("--seed", type=int, help="untrained weights with this seed instead")
    ap.add_argument("--host", default="127.0.0.1")
    ap.add_argument("--port", type=int, default=8501)
    ap.add_argument("--device", type=int, default=0)
    args = ap.parse_args()
    from . import tfrecmodel
    mod = getattr(tfrecmodel, args.model)
    model = mod.load(savedmodel=args.savedmodel, seed=args.seed, device=args.device)
    srv = serve({args.name: (model.spec, model.predict)}, args.host, args.port)
    print("serving %s as /v1/models/%s:predict on %s:%d (%s)"
          % (args.model, args.name, args.host, args.port, model.kernel_name))
    srv.serve_forever()


if __name__ == "__main__":
    main()
