"""Extracts the plot area of the reference's published README figure (assets/location.png: plt.imshow of the
800x800 hit-location image of the quick-start, README.md:24-51) into tests/golden/readme_location_png.npz.
The figure is the only numerical output the reference publishes for this path (known answer K3).
Run where /root/reference exists:  python tests/golden/make_location_fixture.py"""
import os

import numpy as np
from PIL import Image

im = np.array(Image.open("/root/reference/assets/location.png").convert("RGB"))
black = im.sum(axis=2) < 30
rows = np.where(black.sum(axis=1) > 200)[0]
cols = np.where(black.sum(axis=0) > 200)[0]
crop = im[rows.min():rows.max() + 1, cols.min():cols.max() + 1]
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "readme_location_png.npz")
np.savez_compressed(out, plot_area=crop, source="assets/location.png of lcp29/trimesh-ray-optix (matplotlib imshow of locs[800,800,3])")
print("wrote", out, crop.shape, "non-black fraction", float((crop.sum(axis=2) > 40).mean()))
