import sys, time, json
sys.path.insert(0,'/root/repo')
import numpy as np
import tacs_b200
from tacs_b200 import TACS as T, meshgen
lib=tacs_b200.load(); assert lib.init(0)==0
def run(name, mesh, elem, reps=5):
    t0=time.time(); cr,a=meshgen.build_model(T,lib,mesh,[elem]); A=a.createMat(); t1=time.time()
    res,x,y=a.createVec(),a.createVec(),a.createVec()
    x.setArray(meshgen.hash_vector(x.getSize())); a.applyBCs(x); a.setVariables(x)
    ne=a.getNumElements(); bs,nr,nc,nnzb=A.getSizes()
    lib.time_assemble_jacobian(a.h,1.0,0.0,0.0,res.h,A.h,2)
    import ctypes as C
    lib.profile_enable(1); mk=np.zeros(8); ck=np.zeros(8,np.int64)
    lib.profile_collect(mk.ctypes.data_as(C.POINTER(C.c_double)), ck.ctypes.data_as(C.POINTER(C.c_long)))
    ms=lib.time_assemble_jacobian(a.h,1.0,0.0,0.0,res.h,A.h,reps)/reps
    lib.profile_collect(mk.ctypes.data_as(C.POINTER(C.c_double)), ck.ctypes.data_as(C.POINTER(C.c_long))); lib.profile_enable(0)
    kern=dict(element=round(mk[0]/reps,3), gather_res=round(mk[1]/reps,3), gather_mat=round(mk[2]/reps,3), bcs=round(mk[3]/reps,3))
    msr=lib.time_assemble_res(a.h,res.h,reps)/reps
    lib.time_mat_mult(A.h,x.h,y.h,3); mss=lib.time_mat_mult(A.h,x.h,y.h,20)/20
    bytes_=nnzb*(8*bs*bs+4)+4*(nr+1)+16*bs*nr
    print(json.dumps(dict(name=name,elems=ne,nnzb=nnzb,setup_s=round(t1-t0,2),jac_ms=round(ms,3),kernels=kern,jac_elem_per_s=ne/ms*1e3,res_ms=round(msr,3),spmv_ms=round(mss,4),spmv_gbs=bytes_/mss*1e-6, ynorm=y.norm())),flush=True)
run('quad4_300',meshgen.plate(2,300,300),meshgen.iso_shell_element(T,lib,2))
run('quad4_1000',meshgen.plate(2,1000,1000),meshgen.iso_shell_element(T,lib,2))
run('quad9_120',meshgen.plate(3,120,120),meshgen.iso_shell_element(T,lib,3))
run('hex8_40',meshgen.cube(2,40),meshgen.solid_element(T,lib,2))
run('hex8_100',meshgen.cube(2,100),meshgen.solid_element(T,lib,2))
run('hex27_10',meshgen.cube(3,10),meshgen.solid_element(T,lib,3))
run('hex27_30',meshgen.cube(3,30),meshgen.solid_element(T,lib,3))
