"""Large-scene check: P Gaussians (default 12M, R ~ 28M, tiles with > 4096 entries) through ours and the compiled reference;
prints whether radii and images are bit-identical."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autovfx_b200 import scene
from tests import helpers as Hh
dev = torch.device('cuda:0')
P = int(sys.argv[1]) if len(sys.argv) > 1 else 12_000_000
g = scene.config3_scene(P=P, seed=99)
cams = scene.cameras_from_trajectory(scene.trajectory_dict(num_views=300))
for ci in (3, 150):
    a = Hh.resolve(dict(g=g, cam=cams[ci], sh_degree=3, bg=(0.0, 0.1, 0.0), scale_modifier=1.0), dev)
    o = Hh.run_ours(a, debug=False)
    r = Hh.run_ref(a)
    print(P, ci, o['stats'], 'ref R', r['num_rendered'],
          'radii eq', bool(torch.equal(o['radii'], r['radii'])),
          'img eq', [bool(torch.equal(o[k], r[k])) for k in ('color', 'depth', 'alpha')])
    del o, r, a
    torch.cuda.empty_cache()
