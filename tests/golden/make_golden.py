"""Regenerates the committed fixtures from the reference tree (run in the build container, where
/root/reference exists; the GPU box has no reference tree, so tests only read the fixtures).

  suzanne_mesh.npz    examples/assets/suzanne.obj parsed with tobj-0.1.3 indexing (scenes.load_obj):
                      vertices [1966,6] f32 (pos3+normal3), vertices_uv [N,8], indices u32
  suzanne_gold_500.npz examples/suzanne.png (2000x2000 RGBA8) box-downsampled 4x to 500x500 RGB, uint8
"""
import os
import sys

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from softrender_b200 import scenes  # noqa: E402

REF = "/root/reference"
m = scenes.load_obj(os.path.join(REF, "examples/assets/suzanne.obj"))
muv = scenes.load_obj(os.path.join(REF, "examples/assets/suzanne.obj"), with_uv=True)
np.savez_compressed(os.path.join(HERE, "suzanne_mesh.npz"), vertices=m.vertices, indices=m.indices,
                    vertices_uv=muv.vertices, indices_uv=muv.indices)
gold = np.asarray(Image.open(os.path.join(REF, "examples/suzanne.png")).convert("RGB")).astype(np.float32)
g = gold.reshape(500, 4, 500, 4, 3).mean(axis=(1, 3))
np.savez_compressed(os.path.join(HERE, "suzanne_gold_500.npz"), rgb=np.round(g).astype(np.uint8))
print("vertices", m.vertices.shape, "triangles", m.ntris, "uv-vertices", muv.vertices.shape)
