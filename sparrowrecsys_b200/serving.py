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
numbers, "" for strings), scored, and returned in request order.  Errors use TF-Serving's
shape: HTTP 400 `{"error": "..."}`.

Two things sit between the socket and the library call:

* **feature fill** - with `--features <samples.csv>` (or a `FeatureStore` passed to
  `serve`), model inputs that no instance of a request carries are read from the
  reference's `uf:<userId>` / `mf:<movieId>` hashes (`featurestore.py`), so the Java
  server, which posts only `(userId, movieId)`, can be answered by DIN / DeepFM / W&D;
* **cross-request micro-batching** (`MicroBatcher`) - Jetty runs one blocking POST per
  worker thread (`online/util/HttpClient.java:21-40`); requests that arrive while a
  library call is in flight are concatenated into the next call and the scores split back
  per request, so the GPU sees one batch per round trip instead of one per thread.
"""
from __future__ import annotations

import argparse
import json
import queue
import re
import threading
import time
from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer
from typing import Callable, Dict, List, Mapping, Optional

import numpy as np

from .spec import MOVIE_GENRE_KEYS, USER_GENRE_KEYS, ModelSpec

_PATH = re.compile(r"^/v1/models/([^/:]+)(?:/versions/\d+)?:predict$")
_STRING_KEYS = set(MOVIE_GENRE_KEYS) | set(USER_GENRE_KEYS)


def inputs_to_instances(inputs) -> List[Mapping[str, object]]:
    """TF-Serving's columnar request format (`{"inputs": {"userId": [..], "movieId": [..]}}`,
    which TF-Serving answers with `{"outputs": ...}`) -> the row format.  The Java server only
    uses the row format; the columnar one is accepted so that other TF-Serving clients work."""
    if not isinstance(inputs, Mapping) or not inputs:
        raise ValueError("'inputs' must be an object of equally long lists")
    cols = {k: (v if isinstance(v, list) else [v]) for k, v in inputs.items()}
    n = {len(v) for v in cols.values()}
    if len(n) != 1:
        raise ValueError("'inputs' columns differ in length")
    unwrap = lambda x: x[0] if isinstance(x, list) and len(x) == 1 else x      # [[1],[2]] and [1,2]
    return [{k: unwrap(v[i]) for k, v in cols.items()} for i in range(n.pop())]


def instances_to_features(spec: ModelSpec, instances: List[Mapping[str, object]],
                          store=None) -> Dict[str, np.ndarray]:
    """TF-Serving row format -> dict of columns for `predict`.  With a `FeatureStore`,
    inputs that no instance carries come from the `uf:` / `mf:` hashes."""
    if not isinstance(instances, list) or not instances:
        raise ValueError("'instances' must be a non-empty list")
    if not all(isinstance(inst, Mapping) for inst in instances):
        raise ValueError("each instance must be a JSON object (row format)")
    feats: Dict[str, np.ndarray] = {}
    absent = []
    for key in spec.required_keys():
        if store is not None and not any(key in inst for inst in instances):
            absent.append(key)
            continue
        if key in _STRING_KEYS:
            col = np.array([str(inst.get(key, "")) for inst in instances], dtype=object)
        else:
            col = np.array([float(inst.get(key, 0) or 0) for inst in instances], dtype=np.float64)
            if key.endswith("Id") or key.startswith("userRatedMovie"):
                col = col.astype(np.int64)
            else:
                col = col.astype(np.float32)
        feats[key] = col
    if absent:
        from .featurestore import fill_instances
        fill_instances(feats, absent, store, spec.hist_len)
    return feats


class MicroBatcher:
    """Concatenate concurrent requests into one `predict` call.

    `submit(features)` blocks the calling (HTTP worker) thread and returns that request's
    float32 [n,1] scores.  One dispatcher thread takes the oldest waiting request, adds
    whatever else is already queued (up to `max_rows` rows; optionally lingering
    `max_wait_s` for more), scores the concatenation once and splits the result.  With
    `max_wait_s = 0` batching adds no latency: a lone request is dispatched at once, and
    requests pile up only while the previous call is running.  If a merged call fails
    (e.g. one request holds an out-of-range id) its requests are re-run one by one, so an
    error reaches only the request that caused it."""

    def __init__(self, predict_fn: Callable[[Dict[str, np.ndarray]], np.ndarray],
                 max_rows: int = 8192, max_wait_s: float = 0.0):
        self.predict_fn = predict_fn
        self.max_rows = int(max_rows)
        self.max_wait_s = float(max_wait_s)
        self.calls = 0                    # library calls made
        self.requests = 0                 # requests served
        self._q: "queue.Queue" = queue.Queue()
        self._held = None                 # a request taken off the queue that did not fit
        self._stop = False
        self._gate = threading.Lock()     # orders submit's enqueue against close's sentinel
        self._thread = threading.Thread(target=self._run, name="srs-microbatcher", daemon=True)
        self._thread.start()

    def submit(self, feats: Dict[str, np.ndarray]) -> np.ndarray:
        item = {"feats": feats, "n": len(next(iter(feats.values()))), "done": threading.Event(),
                "out": None, "err": None}
        with self._gate:                  # either the request is queued before the close sentinel
            if self._stop:                # (the dispatcher drains it), or it sees the closed flag
                raise RuntimeError("batcher is closed")
            self._q.put(item)
        while not item["done"].wait(timeout=1.0):
            if not self._thread.is_alive():      # dispatcher gone without completing this request
                raise RuntimeError("batcher is closed")
        if item["err"] is not None:
            raise item["err"]
        return item["out"]

    def close(self):
        with self._gate:
            self._stop = True
            self._q.put(None)
        self._thread.join(timeout=5)

    # ---- dispatcher thread ----------------------------------------------------------------
    _EMPTY = object()

    def _take(self, timeout):
        """Next waiting request, `None` (the close sentinel) or `_EMPTY`."""
        if self._held is not None:
            item, self._held = self._held, None
            return item
        try:
            return self._q.get(timeout=timeout) if timeout > 0 else self._q.get_nowait()
        except queue.Empty:
            return self._EMPTY

    def _run(self):
        closing = False
        while not closing:
            first = self._held if self._held is not None else self._q.get()
            self._held = None
            if first is None:
                break
            group, rows = [first], first["n"]
            deadline = time.monotonic() + self.max_wait_s
            while rows < self.max_rows:
                nxt = self._take(deadline - time.monotonic())
                if nxt is None:
                    closing = True
                    break
                if nxt is self._EMPTY:
                    if time.monotonic() >= deadline:
                        break
                    continue
                if rows + nxt["n"] > self.max_rows:
                    self._held = nxt
                    break
                group.append(nxt)
                rows += nxt["n"]
            self._score(group)
        leftovers = [] if self._held is None else [self._held]
        while True:                       # closed: fail whatever is still waiting
            try:
                leftovers.append(self._q.get_nowait())
            except queue.Empty:
                break
        for item in leftovers:
            if item is not None:
                item["err"] = RuntimeError("batcher is closed")
                item["done"].set()

    def _score(self, group):
        try:
            if len(group) == 1:
                merged = group[0]["feats"]
            else:
                merged = {k: np.concatenate([g["feats"][k] for g in group])
                          for k in group[0]["feats"]}
            self.calls += 1
            out = np.asarray(self.predict_fn(merged), np.float32).reshape(-1, 1)
            lo = 0
            for g in group:
                g["out"] = out[lo:lo + g["n"]]
                lo += g["n"]
        except Exception as e:            # noqa: BLE001 - routed to the request(s) below
            if len(group) == 1:
                group[0]["err"] = e
            else:
                for g in group:
                    self._score([g])
                return
        self.requests += len(group)
        for g in group:
            g["done"].set()


def make_handler(models: Mapping[str, tuple], lock: threading.Lock, store=None,
                 batchers: Optional[Mapping[str, MicroBatcher]] = None):
    """`models`: name -> (spec, predict_fn); predict_fn(features) -> float32 [N,1].
    `batchers`: name -> MicroBatcher over that predict_fn (else calls serialise on `lock`)."""

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
                if not isinstance(req, dict):
                    return self._send(400, {"error": "the request body must be a JSON object with "
                                                     "\"instances\" or \"inputs\""})
                columnar = "instances" not in req and "inputs" in req
                feats = instances_to_features(
                    spec, inputs_to_instances(req["inputs"]) if columnar else req.get("instances"), store)
                if batchers and m.group(1) in batchers:
                    p = batchers[m.group(1)].submit(feats)
                else:
                    with lock:                   # one library call at a time per process
                        p = predict_fn(feats)
                scores = [[float(v)] for v in np.asarray(p).reshape(-1)]
                self._send(200, {"outputs": scores} if columnar else {"predictions": scores})
            except (ValueError, KeyError, TypeError, json.JSONDecodeError) as e:
                self._send(400, {"error": str(e)})
            except Exception as e:               # library / CUDA failure, closed batcher: TF-Serving answers
                self._send(500, {"error": "%s: %s" % (type(e).__name__, e)})   # every failure with a JSON error

        def do_GET(self):
            m = re.match(r"^/v1/models/([^/:]+)$", self.path)
            if m and m.group(1) in models:
                return self._send(200, {"model_version_status": [
                    {"version": "1", "state": "AVAILABLE", "status": {"error_code": "OK", "error_message": ""}}]})
            self._send(404, {"error": "not found"})

    return Handler


def serve(models: Mapping[str, tuple], host: str = "127.0.0.1", port: int = 8501, store=None,
          micro_batch: bool = True, max_rows: int = 8192,
          max_wait_s: float = 0.0) -> ThreadingHTTPServer:
    """Build the server (call `.serve_forever()`; `.shutdown()` from another thread stops
    it).  `store`: a `FeatureStore` for inputs the requests do not carry.  `micro_batch`:
    merge concurrent requests per model (`srv.batchers[name]` exposes the counters)."""
    batchers = {name: MicroBatcher(fn, max_rows, max_wait_s) for name, (_, fn) in models.items()} \
        if micro_batch else {}
    srv = _Server((host, port), make_handler(models, threading.Lock(), store, batchers))
    srv.daemon_threads = True
    srv.batchers = batchers
    return srv


class _Server(ThreadingHTTPServer):
    """`server_close()` also stops the per-model dispatcher threads."""
    batchers: Mapping[str, MicroBatcher] = {}

    def server_close(self):
        super().server_close()
        for b in self.batchers.values():
            b.close()


def main():
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--model", default="neuralcf")
    ap.add_argument("--name", default="recmodel", help="servable name in the URL")
    ap.add_argument("--savedmodel", help="reference SavedModel dir (neuralcf / twotowers)")
    ap.add_argument("--seed", type=int, help="untrained weights with this seed instead")
    ap.add_argument("--host", default="127.0.0.1")
    ap.add_argument("--port", type=int, default=8501)
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--features", help="sample CSV (trainingSamples.csv layout) to build the "
                    "uf:/mf: feature store from, for requests that carry only userId/movieId")
    ap.add_argument("--no-micro-batch", action="store_true")
    ap.add_argument("--max-wait-us", type=float, default=0.0,
                    help="linger this long for more requests before a library call")
    args = ap.parse_args()
    from . import tfrecmodel
    mod = getattr(tfrecmodel, args.model)
    model = mod.load(savedmodel=args.savedmodel, seed=args.seed, device=args.device)
    store = None
    if args.features:
        from .featurestore import FeatureStore
        store = FeatureStore.from_samples(args.features)
    srv = serve({args.name: (model.spec, model.predict)}, args.host, args.port, store=store,
                micro_batch=not args.no_micro_batch, max_wait_s=args.max_wait_us * 1e-6)
    print("serving %s as /v1/models/%s:predict on %s:%d (%s)"
          % (args.model, args.name, args.host, args.port, model.kernel_name))
    srv.serve_forever()


if __name__ == "__main__":
    main()
