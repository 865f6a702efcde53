"""Generates the committed golden fixtures from the reference tree (run in the authoring container,
where /root/reference exists; the GPU box and the driver never need it):

    python tests/golden/make_golden.py

Outputs (small, committed):
  out_single_epoch_probe.json  pixels of report/out_single_epoch.png — the reference's own render of
                               the deterministic pass (main.rs:1086-1115) after post_process + sRGB/u8 —
                               on a 40x30 lattice plus SURVEY's hand-picked probes, with the image sha256.
  dodeca_mesh.json             the 20 `v` and 36 `f` statements of dodecahedron.obj as token lists, so
                               the OBJ importer can be exercised without the reference tree.
"""
import hashlib
import json
import os

import numpy as np
from PIL import Image

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    path = os.path.join(REF, "report", "out_single_epoch.png")
    raw = open(path, "rb").read()
    img = np.array(Image.open(path).convert("RGB"))
    h, w, _ = img.shape
    probes = []
    for y in range(16, h, 32):
        for x in range(16, w, 32):
            probes.append([int(y), int(x)] + [int(c) for c in img[y, x]])
    for (y, x) in [(0, 0), (480, 640), (100, 700), (300, 300), (100, 1100), (50, 720), (900, 1100)]:
        probes.append([y, x] + [int(c) for c in img[y, x]])
    out = {
        "source": "report/out_single_epoch.png",
        "sha256": hashlib.sha256(raw).hexdigest(),
        "width": int(w), "height": int(h),
        "black_pixels": int((img.max(axis=2) == 0).sum()),
        "saturated_pixels": int((img.max(axis=2) == 255).sum()),
        "mean_rgb": [float(v) for v in img.reshape(-1, 3).mean(axis=0)],
        "probes_y_x_r_g_b": probes,
    }
    with open(os.path.join(HERE, "out_single_epoch_probe.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))

    v, faces = [], []
    for line in open(os.path.join(REF, "dodecahedron.obj")):
        t = line.split()
        if not t:
            continue
        if t[0] == "v":
            v.append(t[1:4])
        elif t[0] == "f":
            faces.append([int(i) for i in t[1:]])
    with open(os.path.join(HERE, "dodeca_mesh.json"), "w") as f:
        json.dump({"source": "dodecahedron.obj", "v": v, "f": faces}, f, separators=(",", ":"))
    print(len(probes), "probes;", len(v), "vertices;", len(faces), "faces")


if __name__ == "__main__":
    main()
