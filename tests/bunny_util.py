"""The reduced `bunny` fixture (tests/golden/bunny, made by tests/golden/make_bunny.py from the
reference's example/project.xml and example/images/bunny): 8 calibrated, lens-distorted cameras
and 256x192 RGBA images (alpha = object mask), used at image_scale 0.25."""
import json
import os

import numpy as np

from stereoreconstruction_b200 import types as T

DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bunny")


def load(refractive=False):
    """cams (SrCamera list, Camera::setP semantics), rgba images, masks (255 where alpha == 255), scale.
    refractive=True injects the interface BASELINE/SURVEY cfg1 names: (px,py) = principal point +
    (40,25) px, dist = 10, n = 1.333 — the GUI does the same from its spinners
    (gui/widgets/stereowidget.cpp:531-536); example/project.xml itself has no interface."""
    from PIL import Image
    meta = json.load(open(os.path.join(DIR, "cameras.json")))
    cams, imgs, masks = [], [], []
    for c in meta["cameras"]:
        cam = T.camera_from_P(c["P"], dist=c["dist"])
        if refractive:
            T.set_interface_px(cam, cam.K[2] + 40.0, cam.K[5] + 25.0, 10.0, 1.333)
        cams.append(cam)
        im = np.asarray(Image.open(os.path.join(DIR, c["id"] + ".png")).convert("RGBA")).copy()
        masks.append(np.where(im[..., 3] == 255, 255, 0).astype(np.uint8))
        imgs.append(im)
    return cams, imgs, masks, float(meta["scale"])
