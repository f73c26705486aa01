"""where does the time of picstep_step_host go (run on the GPU box)"""
import sys, os, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from picongpu_b200 import param as prm, picstep
p = prm.khi_params(grid=(256, 256, 256))
s = picstep.Simulation(p, device=0, exact=False)
s.init_khi(); s.step(3); s.sync()
sp = []
for name in ("e", "i"):
    n = s.particle_count(name)
    arrs = [torch.empty((3, n), dtype=torch.float32).pin_memory(), torch.empty((3, n), dtype=torch.float32).pin_memory(),
            torch.empty((n,), dtype=torch.float32).pin_memory(), torch.empty((n,), dtype=torch.int32).pin_memory()]
    s.download_particles(name, out=tuple(t.numpy() for t in arrs))
    sp.append(tuple(t.numpy() for t in arrs))
E = torch.from_numpy(s.download_field(0)).pin_memory().numpy(); B = torch.from_numpy(s.download_field(1)).pin_memory().numpy()
def T(f, *a):
    s.sync(); t0 = time.perf_counter(); f(*a); s.sync(); return (time.perf_counter() - t0) * 1e3
for rep in range(2):
    print("upload e %.1f ms, upload i %.1f ms, fields up %.1f ms, step %.1f ms, fields down %.1f ms, energies %.1f ms, step_host %.1f ms" % (
        T(s.upload_particles, "e", *sp[0]), T(s.upload_particles, "i", *sp[1]), T(lambda: (s.upload_field(0, E), s.upload_field(1, B))),
        T(s.step, 1), T(lambda: (s.download_field(0), s.download_field(1))), T(lambda: (s.field_energy(), s.particle_energy("e"), s.particle_energy("i"))),
        T(s.step_host, E, B, sp)))
n = sp[0][2].shape[0]
print("bytes per species %.2f GB" % (n * 32 / 1e9))
