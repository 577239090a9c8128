#!/usr/bin/env python3
"""Sweep rows-per-thread of the y-marching stencil kernels (env T3D_*_ROWS is read once per process)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, torch; sys.path.insert(0, %r)
import bench
from tomography_3d_reconstructor_b200 import engine
Z,H,W=512,1024,1024
dev=torch.device("cuda",0)
masks=bench.make_phantom_u8(Z,H,W,0,Z,dev)
dv=engine.pack_and_close(masks,200,True)
def t(fn,n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b)/n*1e3
sm=engine.smooth(dv,3,True)
print("morph4x %%.1f us  sign_lean %%.1f us" %% (t(lambda: engine.smooth(dv,3,True)), t(lambda: engine.field_sign(sm,1,lean=True))))
sign,dims,_,_=engine.field_sign(sm,1,lean=True)
L=engine._L(); n=int(L.t3d_mc_num_chunks(*dims)); bal=torch.empty(n,dtype=torch.int32,device=dev)
print("flags %%.1f us" %% t(lambda: engine.check(L.t3d_mc_flags(engine._p(sign),*dims,0,-1,engine._p(bal),engine._stream()))))
''' % ROOT
for rows in (4, 8, 16, 32, 64):
    env = dict(os.environ, T3D_MORPH_ROWS=str(rows), T3D_SIGN_ROWS=str(rows), T3D_FLAGS_ROWS=str(rows))
    out = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True)
    print("rows", rows, out.stdout.strip().replace("\n", " | "), out.stderr.strip()[-300:])
