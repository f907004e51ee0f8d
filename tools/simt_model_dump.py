"""Writes /tmp/sim/layout.bin and /tmp/sim/{prim,b1,b2,b3}.bin for tools/simt_model.cpp (development aid)."""
import os
os.makedirs("/tmp/sim", exist_ok=True)
# dumps mode-2 layouts (8 octants) and a set of rays (camera + bounce1 + bounce2) for the SIMT model
import sys, importlib, numpy as np, ctypes as C
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
pkg=importlib.import_module("zig-raytracing-weekend_b200")
import oracle_ffi as orc
world=pkg.World.book1()
lib=pkg._ffi.rtb()
n=C.c_uint32(0)
lib.rtb_debug_build_layout(world.desc,2,0,None,C.byref(n))
L=np.zeros((8,n.value+1,8),np.float32)
for o in range(8):
    assert lib.rtb_debug_build_layout(world.desc,2,o,L[o].ctypes.data,C.byref(n))==0
L.tofile('/tmp/sim/layout.bin'); print("nodes",n.value)
cam=pkg.book1_camera(1200,500,50).init()
rng=np.random.default_rng(1)
# camera rays in tile order: pick random 32x... use 8x4 pixel blocks like the kernel's warps
NW=4000
rays=[]
W,H=1200,675
for w in range(NW):
    x0=rng.integers(0,W-8); y0=rng.integers(0,H-4)
    px=[(y0+dy)*W+(x0+dx) for dy in range(4) for dx in range(8)]
    rays.append(orc.get_rays(cam,7,np.array(px),np.full(32,w%64)))
prim=np.concatenate(rays)
def bounce(rs,seg):
    h=orc.trace_rays(world.desc,rs)
    out=[]
    for i in np.nonzero(h["object"]>=0)[0]:
        ok,_,sc=orc.scatter(world.desc,rs[i:i+1],h[i:i+1],5,int(i),0,seg)
        if ok: out.append(sc)
    return np.concatenate(out)
b1=bounce(prim,1); b2=bounce(b1,2); b3=bounce(b2,3)
print(len(prim),len(b1),len(b2),len(b3))
def save(name,rs,shuffle):
    a=np.zeros((len(rs),8),np.float32)
    a[:,0:3]=rs["origin"]; a[:,3:6]=rs["direction"]; a[:,6]=rs["time"]
    if shuffle: a=a[rng.permutation(len(a))]
    a.tofile(f'/tmp/sim/{name}.bin')
save("prim",prim,False); save("b1",b1,True); save("b2",b2,True); save("b3",b3,True)
