#include <cstdio>
__device__ __forceinline__ int imin(int a,int b){int r; asm("min.s32 %0, %1, %2;":"=r"(r):"r"(a),"r"(b)); return r;}
__device__ __forceinline__ int imax(int a,int b){int r; asm("max.s32 %0, %1, %2;":"=r"(r):"r"(a),"r"(b)); return r;}
__device__ __forceinline__ int sel_min(int a,int b){return a<b?a:b;}
__device__ __forceinline__ int sel_max(int a,int b){return a>b?a:b;}
template<int V> __device__ __noinline__ int best_v(const int* d){ int best=-1000; for(int s=0;s<16;s++){int mn=d[s],mx=d[s];
  for(int k=1;k<9;k++){int x=d[(s+k)&15];
    if(V==0){mn=min(mn,x);mx=max(mx,x);} if(V==1){mn=imin(mn,x);mx=imax(mx,x);} if(V==2){mn=sel_min(mn,x);mx=sel_max(mx,x);} }
  int t; if(V==0) t=max(mn,-mx); if(V==1) t=imax(mn,-mx); if(V==2) t=sel_max(mn,-mx);
  if(V==0) best=max(best,t); if(V==1) best=imax(best,t); if(V==2) best=sel_max(best,t);
  if(V==3){ /* separate: bright and dark passes */ }
 }return best;}
// two separate reductions: dark = max over arcs of min d ; bright = max over arcs of min(-d)
__device__ __noinline__ int best_sep(const int* d){ int bd=-1000,bb=-1000; for(int s=0;s<16;s++){int mn=d[s],mn2=-d[s];for(int k=1;k<9;k++){int x=d[(s+k)&15];mn=min(mn,x);mn2=min(mn2,-x);}bd=max(bd,mn);bb=max(bb,mn2);} return max(bd,bb);}
__global__ void dbg(const int* d_in, int* out){ int d[16]; for(int k=0;k<16;k++) d[k]=d_in[k]; out[0]=best_v<0>(d); out[1]=best_v<1>(d); out[2]=best_v<2>(d); out[3]=best_sep(d);
  // probe single ops
  out[4]=max(27,-37); out[5]=max(d[5],-d[15]); int mn=d[14]; for(int k=15;k<23;k++) mn=min(mn,d[k&15]); out[6]=mn; int mx=d[14]; for(int k=15;k<23;k++) mx=max(mx,d[k&15]); out[7]=mx; out[8]=max(mn,-mx); }
int main(){ int h[16]={33,32,30,29,31,27,30,20,-1,-1,0,-2,-2,7,36,37}; int *di,*dout; cudaMalloc(&di,64); cudaMalloc(&dout,256);
 cudaMemcpy(di,h,64,cudaMemcpyHostToDevice); dbg<<<1,1>>>(di,dout); int o[32]; cudaMemcpy(o,dout,128,cudaMemcpyDeviceToHost);
 printf("v0=%d v1=%d v2=%d sep=%d (expect 27) | probes %d %d mn=%d mx=%d t=%d (expect 27 27 27 37 27)\n",o[0],o[1],o[2],o[3],o[4],o[5],o[6],o[7],o[8]); return 0; }
