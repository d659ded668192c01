"""Regenerate the committed golden fixtures from the reference checkout.

Run in the build container (needs /root/reference; the GPU box has none):

    python tests/golden/make_golden.py

Writes, next to this file:

* `samples_head.csv`    - header + first 512 data rows of the reference's
  `src/main/resources/webroot/sampledata/testSamples.csv`, byte-for-byte.
* `neuralcf_002.npz`, `neuralcf_001.npz`, `mlprec_005.npz` - the reference's shipped
  trained weights (`modeldata/neuralcf/{002,001}`, `modeldata/MLPRec/005`) read with
  `sparrowrecsys_b200.bundle` (no TensorFlow): Dense kernels/biases and the movie
  table in full, the 30001-row user table only for the users that occur in
  `samples_head.csv` plus userId 10351 (the pair `HttpClient.main` posts,
  `online/util/HttpClient.java:110-147`); `user_ids` / `user_rows` rebuild a
  zero-filled table.
* `item2vecEmb.csv`, `userEmb_head.csv` - the reference's shipped embeddings for the "emb"
  ranker (`modeldata/item2vecEmb.csv` byte-for-byte; first 20 lines of `modeldata/userEmb.csv`),
  in the `id:f f f ...` format `DataManager.loadMovieEmb` / `Utility.parseEmbStr` read.
* `full_file_stats.json` - accuracy / ROC-AUC of neuralcf/002 over all 22 440 test
  rows (oracle output), for the whole-file sanity check quoted in SURVEY.md 8c.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from sparrowrecsys_b200 import bundle, features                      # noqa: E402
from sparrowrecsys_b200.spec import default_spec                      # noqa: E402
from oracle import ctr_oracle                                        # noqa: E402

REF = "/root/reference/src/main/resources/webroot/"
N_HEAD = 512


def copy_embeddings():
    with open(REF + "modeldata/item2vecEmb.csv", "rb") as f:
        data = f.read()
    with open(os.path.join(HERE, "item2vecEmb.csv"), "wb") as f:
        f.write(data)
    with open(REF + "modeldata/userEmb.csv", "rb") as f:
        head = b"".join(f.readline() for _ in range(20))
    with open(os.path.join(HERE, "userEmb_head.csv"), "wb") as f:
        f.write(head)


def main():
    copy_embeddings()
    src = REF + "sampledata/testSamples.csv"
    with open(src, "rb") as f:
        lines = f.read().split(b"\n")
    with open(os.path.join(HERE, "samples_head.csv"), "wb") as f:
        f.write(b"\n".join(lines[:N_HEAD + 1]) + b"\n")
    head = features.load_samples_csv(os.path.join(HERE, "samples_head.csv"))
    users = np.unique(np.concatenate([head["userId"], np.array([10351], np.int32)]))
    for name, loader, path in (("neuralcf_002", bundle.load_neuralcf, "modeldata/neuralcf/002"),
                               ("neuralcf_001", bundle.load_neuralcf, "modeldata/neuralcf/001"),
                               ("mlprec_005", bundle.load_twotowers, "modeldata/MLPRec/005")):
        W = loader(REF + path)
        out = {k.replace("/", "__"): v for k, v in W.items() if k != "userId_embedding"}
        out["user_ids"] = users.astype(np.int32)
        out["user_rows"] = W["userId_embedding"][users]
        np.savez(os.path.join(HERE, name + ".npz"), **out)
    full = features.load_samples_csv(src)
    W = bundle.load_neuralcf(REF + "modeldata/neuralcf/002")
    p = ctr_oracle.predict(default_spec("neuralcf"), W, full)[:, 0]
    lab = full["label"]
    order = np.argsort(p, kind="mergesort")
    ranks = np.empty(len(p)); ranks[order] = np.arange(1, len(p) + 1)
    # average ranks over ties
    _, inv, cnt = np.unique(p, return_inverse=True, return_counts=True)
    sums = np.bincount(inv, weights=ranks)
    ranks = (sums / cnt)[inv]
    npos = int((lab == 1).sum()); nneg = len(lab) - npos
    auc = (ranks[lab == 1].sum() - npos * (npos + 1) / 2) / (npos * nneg)
    stats = {"rows": int(len(p)), "accuracy": float(((p > 0.5) == (lab == 1)).mean()),
             "roc_auc": float(auc)}
    with open(os.path.join(HERE, "full_file_stats.json"), "w") as f:
        json.dump(stats, f, indent=1)
    print(stats)


if __name__ == "__main__":
    main()
