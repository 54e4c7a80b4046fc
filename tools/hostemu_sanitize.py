"""The device code under AddressSanitizer + UndefinedBehaviorSanitizer (no GPU needed): builds tests/hostemu with
-fsanitize=address,undefined and runs the device headers (tile / step / stats paths, ray streams, the divergence model)
and, with the argument `kernels`, every __global__ kernel on the coroutine SIMT emulator -- shared-memory queues,
record planes, stack indexing, shifts.  ASan only partly follows swapcontext; it reports nothing on this code.

  python tools/hostemu_sanitize.py [kernels]      (re-executes itself with libasan preloaded)
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = "/tmp/libsvo_hostemu_asan.so"
if os.environ.get("SVO_SANITIZE_CHILD") != "1":
    he, cs = os.path.join(ROOT, "tests", "hostemu"), os.path.join(ROOT, "svo_raytracer_b200", "csrc")
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
                           "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-Wno-unknown-pragmas",
                           "-I" + os.environ.get("CUDA_HOME", "/usr/local/cuda") + "/include", "-o", LIB] +
                          [os.path.join(he, f) for f in ("emu.cpp", "kernels_emu.cpp", "wavefront_emu.cpp", "gpu_build_emu.cpp", "transcode_emu.cpp", "simt_emu.cpp")] +
                          [os.path.join(cs, "svo_transcode.cpp"), "-lpthread"])
    asan = subprocess.check_output(["gcc", "-print-file-name=libasan.so"]).decode().strip()
    env = dict(os.environ, SVO_SANITIZE_CHILD="1", LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0:detect_stack_use_after_return=0")
    sys.exit(subprocess.call([sys.executable, os.path.abspath(__file__)] + sys.argv[1:], env=env))

sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, ROOT)
from hostemu import emu as E
E._LIB_PATH = LIB
E.build = lambda force=False: E._LIB_PATH
import numpy as np
import svo_raytracer_b200 as svo
from oracle import oracle as O
hm,mm=svo.terrain_inputs(128); nodes,_=O.build_terrain(hm,mm,128,64)
sc=E.Scene(nodes)
W,H=200,120
for cam,mode,casts in (("B",0,2),("C",2,2),("A",0,4)):
    pos,l1,l2,r1,r2=svo.CAMERAS[cam]
    f=O.make_frame(pos,l1,l2,r1,r2,frame_number=1,render_mode=mode,max_depth=7,casts=casts)
    for path in (E.PATH_RUN,E.PATH_STEP,E.PATH_STATS):
        sc.render(f,W,H,path=path)
    sc.render(f,W,H,box=True,aux=False)
    sc.simt(f,W,H,[14,40,18,31,2,8,10,700])
    print("plain paths ok",cam,mode,flush=True)
rng=np.random.default_rng(1)
rays=np.zeros(3000,dtype=O.RAY_DTYPE); rays["o"]=rng.uniform(0.9,2.1,(3000,3)); d=rng.normal(size=(3000,3)); rays["d"]=d/np.linalg.norm(d,axis=1,keepdims=True)
rays["d"][::50]=0; rays["d"][7::60]=np.nan
sc.cast(rays,7); sc.simt_stream(rays,[14,40,18,31,2,8,10,300,230,6.5],max_depth=7)
print("streams ok",flush=True)
if len(sys.argv)>1:
    for k in (0,10,13,9,5,4,6,7,8,15,16,1,2,14):
        pos,l1,l2,r1,r2=svo.CAMERAS["B"]
        f=O.make_frame(pos,l1,l2,r1,r2,frame_number=1,render_mode=0,max_depth=7,casts=3)
        sc.launch_render(f,W,H,kernel=k,aux=True,nthreads=1); sc.launch_render(f,W,H,kernel=k,aux=False,box=True,nthreads=1)
        print("emulated kernel",k,"ok",flush=True)
    for k in (0,1,2):
        sc.launch_cast(rays,7,kernel=k,nthreads=1)
    print("emulated stream kernels ok",flush=True)
    for k in (17,19):
        sc.launch_render(f,W,H,kernel=k,aux=False,box=True,nthreads=1)
    beam=sc.beam_conservative(f,W,H,nthreads=1)
    assert np.array_equal(sc.beam_in_parts(f,W,H,3,nthreads=1).view(np.uint32),beam.view(np.uint32))
    print("tile queue, conservative beam (whole and in parts) ok",flush=True)
    # device world builder, whole and incremental transcode
    hm64,mm64=svo.terrain_inputs(64)
    assert np.array_equal(E.gpu_build_terrain(hm64,mm64,64,32),svo.build_terrain(hm64,mm64,64,32))
    assert E.gpu_transcode_check(nodes)==0
    import svo_stream as S
    ed=S.StreamEditor(nodes)
    p=ed.find_leaf(2,False,rng,min_depth=3)
    ed.subdivide(p,[0,2,0,0,3,0,0,1])
    ed.set_value(ed.find_leaf(1,True,rng),0)
    r=E.patch_check(nodes,ed.stream(),ed.ranges())
    assert r["status"]==0 and r["fell_back"]==0,r
    print("device builder, whole and incremental transcode ok")
