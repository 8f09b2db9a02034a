import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))) if '__file__' in globals() else '/root/repo')
sys.path.insert(0, '/root/repo')
from autovfx_b200 import scene, rasterizer as R
from autovfx_b200 import render_loop as RL
dev = torch.device('cuda:0')
g = {k: v.to(dev) for k, v in scene.config3_scene().items()}
cams = scene.cameras_from_trajectory(scene.trajectory_dict(num_views=300))
packed = RL.pack_cameras(cams).to(dev)
bg = torch.zeros(3, device=dev)
W, H = 1920, 1080
def make(nstreams):
    streams = [torch.cuda.Stream(dev) for _ in range(nstreams)]
    outs = [(torch.empty((3, H, W), device=dev), torch.empty((1, H, W), device=dev), torch.empty((1, H, W), device=dev), torch.empty((3_000_000,), dtype=torch.int32, device=dev)) for _ in range(nstreams)]
    camrows = [torch.empty(37, device=dev) for _ in range(nstreams)]
    pfs = []
    for s in range(nstreams):
        with torch.cuda.stream(streams[s]):
            pfs.append(R.PreparedForward(g['means3D'], g['shs'], g['opacities'], g['scales'], g['rotations'], camrows[s], W, H, bg, 3, 1.0, outs[s]))
    return streams, camrows, pfs
def run(nstreams, K=300):
    streams, camrows, pfs = make(nstreams)
    tf = [(float(packed[i, 35]), float(packed[i, 36])) for i in range(300)]
    def frame(i):
        s = i % nstreams
        with torch.cuda.stream(streams[s]):
            camrows[s].copy_(packed[i % 300], non_blocking=True)
            return pfs[s].launch(*tf[i % 300])
    # size capacity
    for i in range(0, 300, 7):
        t = frame(i); 
        if not t.ok():
            t = frame(i)
    torch.cuda.synchronize()
    for i in range(10): frame(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ts = [frame(i) for i in range(K)]
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    bad = sum(1 for t in ts if t.stats()['overflow'])
    return K / dt, bad
for n in (1, 2, 3, 1, 2):
    print(n, 'streams: %.1f fps, overflow %d' % run(n))
