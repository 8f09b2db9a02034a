"""Drop-in name for the reference package (DGR/diff_gaussian_rasterization/__init__.py): callers such as
sugar/gaussian_splatting/gaussian_renderer/__init__.py:16 and sugar/sugar_scene/sugar_model.py:9 do
``from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer``."""
from autovfx_b200.rasterizer import (GaussianRasterizationSettings, GaussianRasterizer, _RasterizeGaussians,  # noqa: F401
                                     rasterize_gaussians)
