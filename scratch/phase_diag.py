"""Where does the Pkphase difference of the fused (density) path come from?  One grid, four routes."""
import sys
import numpy as np
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from conftest import make_particles, BOX
from oracle import build as ob
ob.build()
from oracle import cpu as O
from pylians3_b200 import MAS_library as MASL, Pk_library as PKL, prebias_

N = 96
pos, W = make_particles(41, 3 * N ** 3, True)
ref = np.zeros((N, N, N), np.float32)
O.MA(pos, ref, BOX, "PCS", W)
dens_o = ref.copy()
ref /= np.mean(ref, dtype=np.float64); ref -= 1.0
want = O.Pk(ref, BOX, 0, "PCS", 1, False)
pos_d, W_d = torch.from_numpy(pos).cuda(), torch.from_numpy(W).cuda()
grid = torch.empty((N, N, N), dtype=torch.float32, device="cuda")
c = prebias_(grid, len(pos), W_d)
MASL.MA(pos_d, grid, BOX, "PCS", W_d)
sel = want.Nmodes3D >= 64


def rep(name, got, base=want):
    r = np.abs(np.asarray(got.Pkphase)[sel] / np.asarray(base.Pkphase)[sel] - 1)
    p = np.abs(np.asarray(got.Pk)[:, 0] / np.asarray(base.Pk)[:, 0] - 1)
    print("%-58s phase max %.2e at shell %d   P0 max %.2e" % (name, r.max(), np.nonzero(sel)[0][r.argmax()], p.max()))


fused = PKL.Pk(grid, BOX, 0, "PCS", verbose=False, density=True, offset=c)
rep("GPU deposit (n-c), fused  vs oracle", fused)
dens_g = (grid.double() + c).float().cpu().numpy()
d_g = dens_g.copy(); d_g /= np.mean(d_g, dtype=np.float64); d_g -= 1.0
classic = PKL.Pk(torch.from_numpy(d_g).cuda(), BOX, 0, "PCS", verbose=False)
rep("GPU deposit -> host delta -> GPU Pk (classic) vs oracle", classic)
rep("fused vs classic (same deposit)", fused, classic)
want_g = O.Pk(d_g, BOX, 0, "PCS", 1, False)
rep("classic GPU Pk vs oracle Pk of the SAME delta", classic, want_g)
rep("fused vs oracle Pk of the GPU deposit's delta", fused, want_g)
c_o = float(np.float32(dens_o.mean(dtype=np.float64)))
fo = PKL.Pk(torch.from_numpy(dens_o - np.float32(c_o)).cuda(), BOX, 0, "PCS", verbose=False, density=True, offset=c_o)
rep("oracle deposit - c, fused on GPU vs oracle", fo)
co = PKL.Pk(torch.from_numpy(ref).cuda(), BOX, 0, "PCS", verbose=False)
rep("oracle delta, classic GPU Pk vs oracle", co)
