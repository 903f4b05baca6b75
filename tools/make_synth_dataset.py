"""Write a tiny synthetic MSR-VTT-style dataset (the layout dataloader.py:354-470 of the reference reads):
<out>/feats/{train,val}/videoN.npy  fp32 [12, 512];  <out>/annotations.json {"videos": [...], "sentences": [...]};
<out>/config.json = the shipped config with paths rewritten, 1 epoch, metric early-stop off (no Java here)."""
import json
import os
import sys

import numpy as np


def make(out: str, reference_root: str, n_train: int = 16, n_val: int = 4, caps_per_video: int = 4, batch: int = 8) -> str:
    rng = np.random.default_rng(0)
    os.makedirs(os.path.join(out, "feats", "train"), exist_ok=True)
    os.makedirs(os.path.join(out, "feats", "val"), exist_ok=True)
    videos, sentences = [], []
    for i in range(n_train + n_val):
        split = "train" if i < n_train else "validate"
        vid = f"video{i}"
        np.save(os.path.join(out, "feats", "train" if split == "train" else "val", vid + ".npy"),
                rng.standard_normal((12, 512)).astype(np.float32))
        videos.append({"video_id": vid, "split": split})
        for _ in range(caps_per_video):
            n = int(rng.integers(3, 12))
            sentences.append({"video_id": vid, "caption": " ".join(f"w{int(t)}" for t in rng.integers(1000, 30522, n))})
    ann = os.path.join(out, "annotations.json")
    with open(ann, "w") as f:
        json.dump({"videos": videos, "sentences": sentences}, f)
    with open(os.path.join(reference_root, "configs", "caption-task_baseline_modal_clip4clip_config.json")) as f:
        cfg = json.load(f)
    for split, sub in (("train", "train"), ("validation", "val"), ("eval", "val")):
        cfg["data"][split]["feat_dir"] = [os.path.join(out, "feats", sub)]
        cfg["data"][split]["annotation_path"] = ann
    cfg["data"]["train"]["batch_size"] = batch
    cfg["data"]["validation"]["batch_size"] = batch
    cfg["train"].update(epoch=1, metric_earlystop=False, save_dir=os.path.join(out, "checkpoint"), log_dir=os.path.join(out, "log"))
    path = os.path.join(out, "config.json")
    with open(path, "w") as f:
        json.dump(cfg, f)
    return path


if __name__ == "__main__":
    print(make(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "/root/reference"))
