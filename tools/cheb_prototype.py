#!/usr/bin/env python
"""CPU prototype (scipy): a HIERARCHY-FREE vertex-space term for meshes without grid structure.
M^-1 r = Binv r + P C_m(A_c) P' r with C_m = m steps of the Jacobi-scaled Chebyshev iteration for A_c = P'AP on the
interval [lambda_max / alpha, lambda_max] (lambda_max from 20 power iterations).  C_m is a fixed polynomial in A_c, so the
preconditioner is linear and SPD and plain CG applies.  Needs only the general (ELL) vertex operator and axpy/SpMV kernels
on the device - no coarsening.  Not part of the product.   python tools/cheb_prototype.py"""
import sys, os, time
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tools')); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import amg_prototype as ap
from mg_prototype import block_jacobi, pcg, prolongation
import argparse
class A_: pass
def run(n=0, delaunay=0, lattice=False, k=1, degs=(8,16,32), alphas=(30,100)):
    a=A_(); a.n=n; a.delaunay=delaunay; a.lattice=lattice; a.seed=3
    mesh=ap.make_mesh(a)
    qd={1:2,2:4,3:6,4:9}[k]
    tab,A,b,isbc=ap.build_system(mesh,k,qd)
    nt=tab.nt
    P,bnode=prolongation(mesh,nt,isbc)
    Ac=(P.T@A@P).tocsr()+sp.diags(bnode.astype(float))
    bj=block_jacobi(A,nt)
    x0,itbj=pcg(A,b,bj,maxit=20000)
    print(f"{mesh.ncells} cells k={k}: block-Jacobi {itbj} its; vertex dofs {Ac.shape[0]}, trace dofs {A.shape[0]}")
    dinv=np.where(bnode,0.0,1.0/Ac.diagonal())
    # lambda_max of D^-1 A_c by power iteration
    v=np.random.default_rng(0).standard_normal(Ac.shape[0]); v[bnode]=0
    for _ in range(20):
        v=dinv*(Ac@v); lam=np.linalg.norm(v); v/=lam
    lmax=1.1*lam
    for alpha in alphas:
        lmin=lmax/alpha
        th,de=(lmax+lmin)/2,(lmax-lmin)/2
        for m in degs:
            def cheb(r):
                # Chebyshev iteration for A_c x = r with Jacobi scaling, m steps, zero start (standard 3-term)
                x=np.zeros_like(r); res=r.copy()
                sig=th/de; rho=1/sig
                d=dinv*res/th
                for i in range(m):
                    x+=d
                    res-=Ac@d
                    rho_new=1/(2*sig-rho)
                    d=rho_new*rho*d+2*rho_new/de*(dinv*res)
                    rho=rho_new
                return x
            M=lambda r: bj(r)+P@cheb(P.T@r)
            x,it=pcg(A,b,M,maxit=5000)
            cost=it*(1+m*Ac.nnz/A.nnz)
            print(f"   alpha {alpha:4d} degree {m:3d}: {it:4d} its, relative cost {cost:7.0f} trace-SpMV equivalents (block-Jacobi: {itbj}), err {np.linalg.norm(x-x0)/np.linalg.norm(x0):.1e}")
run(n=120,k=1)
run(delaunay=25000,lattice=True,k=1,degs=(16,32),alphas=(100,))
