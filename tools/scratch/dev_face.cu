// scratch: host vs device evaluation of the wall integrals (same code, TIT_HD)
#include <cstdio>
#include <vector>
#include <algorithm>
#include "../../oracle/oracle_kernel.h"
#include "../../titsolver_b200/csrc/sph_kernel.cuh"
using namespace titgpu;
__global__ void k(Params P, const FaceFrame<3>* fr, int n, Vec<3> x, double* out_f, double* out_a){
  int i=blockIdx.x*blockDim.x+threadIdx.x; if(i>=n) return;
  out_f[i]=SphKernel<4>::face_integral<false>(P,fr[i],x);
  out_a[i]=SphKernel<4>::face_integral<true>(P,fr[i],x);
}
int main(){
  double h=0.2, dr=0.1;
  Params P{}; P.h=h; P.hinv=1/h; P.tiny=std::pow(2.220446049250313e-16,1.0/3); P.tiny2=P.tiny*P.tiny; P.radius=2*h; P.radius2=P.radius*P.radius;
  double w3=titgpu_gen::KernelGen<4>::weight3; P.w_flux=w3*P.hinv; P.w_anti=w3;
  using OK=orc::Kernel<oracle_gen::K4>;
  double xs[3][3]={{0.5,0.3,0.2},{0.5,0.3,0.0},{0.5,0.0,0.0}};
  for(auto& x: xs){
  std::vector<FaceFrame<3>> frs; std::vector<double> fo, ao;
  for(int i=-2;i<12;i++)for(int j=-2;j<12;j++)for(int t=0;t<2;t++){
    double ax=i*dr, ay=j*dr;
    orc::Triangle T = t==0? orc::Triangle{{ax,ay,0},{ax+dr,ay,0},{ax+dr,ay+dr,0}} : orc::Triangle{{ax,ay,0},{ax+dr,ay+dr,0},{ax,ay+dr,0}};
    orc::Vec<3> X{x[0],x[1],x[2]};
    if(!T.intersects(orc::BSphere<3>{X,2*h})) continue;
    fo.push_back(OK::flux(T,X,h)[2]); ao.push_back(OK::antigrad_flux(T,X,h));
    FaceFrame<3> fr{};
    Vec<3> a{T.a[0],T.a[1],T.a[2]}, b{T.b[0],T.b[1],T.b[2]}, c{T.c[0],T.c[1],T.c[2]};
    Vec<3> ba=b-a, ca=c-a; Vec<3> wn=cross(ba,ca)*0.5; Vec<3> n=normalize(wn,P.tiny2), e1=normalize(ba,P.tiny2), e2=normalize(cross(wn,e1),P.tiny2);
    for(int d=0;d<3;d++){fr.a[d]=a[d];fr.n[d]=n[d];fr.e1[d]=e1[d];fr.e2[d]=e2[d];}
    fr.bx=dot(ba,e1);fr.cx=dot(ca,e1);fr.cy=dot(ca,e2);
    frs.push_back(fr);
  }
  int n=frs.size(); FaceFrame<3>* d_fr; double *d_f,*d_a; cudaMalloc(&d_fr,n*sizeof(FaceFrame<3>)); cudaMalloc(&d_f,n*8); cudaMalloc(&d_a,n*8);
  cudaMemcpy(d_fr,frs.data(),n*sizeof(FaceFrame<3>),cudaMemcpyHostToDevice);
  Vec<3> xx{x[0],x[1],x[2]};
  k<<<(n+63)/64,64>>>(P,d_fr,n,xx,d_f,d_a);
  std::vector<double> gf(n), ga(n); cudaMemcpy(gf.data(),d_f,n*8,cudaMemcpyDeviceToHost); cudaMemcpy(ga.data(),d_a,n*8,cudaMemcpyDeviceToHost);
  printf("x=(%g,%g,%g) n=%d err=%s\n",x[0],x[1],x[2],n,cudaGetErrorString(cudaGetLastError()));
  double so=0,sg=0,sh=0;
  for(int i=0;i<n;i++){
    double hf=SphKernel<4>::face_integral<false>(P,frs[i],xx), ha=SphKernel<4>::face_integral<true>(P,frs[i],xx);
    so+=fo[i]; sg+=gf[i]; sh+=hf;
    if(fabs(gf[i]-fo[i])>1e-11||fabs(ga[i]-ao[i])>1e-11) printf(" face %d a=(%g,%g) bx=%g cx=%g cy=%g flux: oracle %.15g hostprod %.15g dev %.15g | anti: oracle %.15g hostprod %.15g dev %.15g\n",i,frs[i].a[0],frs[i].a[1],frs[i].bx,frs[i].cx,frs[i].cy,fo[i],hf,gf[i],ao[i],ha,ga[i]);
  }
  printf(" sums flux oracle %.15g hostprod %.15g dev %.15g\n",so,sh,sg);
  }
}
