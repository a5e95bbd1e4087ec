import sys; sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo')
import common, numpy as np
pkg = common.pkg
emu = common.emu_backend()
cu = pkg.Backend(); cu.set_table_path(common.table_dir())
ref = common.ref_backend()
inputs = common.make_inputs()
for be in (emu, cu, ref):
    be.state.init(inputs, broadcast_inputs=True, ps=True, sigma=True)
for M in [1e3, 1e8, 5e8*1.0, 1e12, 1e16]:
    print(M, *["%.17g" % be.lib.sigma_z0(M) for be in (emu,cu,ref)])
    print(M, *["%.17g" % be.lib.dsigmasqdm_z0(M) for be in (emu,cu,ref)])
for k in [1e-3, 0.1, 10.]:
    print(k, *["%.17g" % be.lib.power_in_k(k) for be in (emu,cu,ref)])
for z in [8.0, 300.]:
    print(z, *["%.17g" % be.lib.dicke(z) for be in (emu,cu,ref)])
