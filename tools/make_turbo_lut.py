"""Prints the hex dump of OpenCV's TURBO colormap table (B,G,R per entry) embedded in autovfx_b200/renderer.py."""
import textwrap

import cv2
import numpy as np

lut = cv2.applyColorMap(np.arange(256, dtype=np.uint8).reshape(-1, 1), cv2.COLORMAP_TURBO).reshape(256, 3)
print("\n".join(textwrap.wrap(lut.tobytes().hex(), 120)))
