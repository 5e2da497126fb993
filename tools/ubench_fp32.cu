// Microbenchmarks that size the design of the brute-force matching kernel on B200 (sm_100a):
//   (1) register-resident FFMA / FFMA2 / FADD2 peak (the FP32 roofline denominator),
//   (2) the packed matching inner loop at several sources-per-thread / block-size choices.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_fp32 ubench_fp32.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)
__device__ __forceinline__ u64 pack2(float a, float b){ u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack2(u64 v, float&a, float&b){ asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 sub2(u64 a, u64 b){ u64 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b){ u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c){ u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float min3(float a, float b, float c){ float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

// ---- (1) peak kernels: CH independent chains per thread, ITER iterations
template<int CH> __global__ void __launch_bounds__(256) peak_ffma(float* out, int iters, float a, float b){
  float v[CH];
  #pragma unroll
  for(int c=0;c<CH;c++) v[c]=threadIdx.x*0.001f+c;
  for(int i=0;i<iters;i++){
    #pragma unroll
    for(int c=0;c<CH;c++) v[c]=fmaf(v[c],a,b);
  }
  float s=0; 
  #pragma unroll
  for(int c=0;c<CH;c++) s+=v[c];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int CH> __global__ void __launch_bounds__(256) peak_ffma2(float* out, int iters, float a, float b){
  u64 v[CH]; u64 A=pack2(a,a*1.0001f), B=pack2(b,b*1.0001f);
  #pragma unroll
  for(int c=0;c<CH;c++) v[c]=pack2(threadIdx.x*0.001f+c, threadIdx.x*0.002f+c);
  for(int i=0;i<iters;i++){
    #pragma unroll
    for(int c=0;c<CH;c++) v[c]=fma2(v[c],A,B);
  }
  float s=0; 
  #pragma unroll
  for(int c=0;c<CH;c++){ float x,y; unpack2(v[c],x,y); s+=x+y; }
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
// FADD2 with scalar-broadcast first operand (the form the matching loop uses)
template<int CH> __global__ void __launch_bounds__(256) peak_fadd2(float* out, int iters, float a){
  u64 v[CH];
  #pragma unroll
  for(int c=0;c<CH;c++) v[c]=pack2(threadIdx.x*0.001f+c, threadIdx.x*0.002f+c);
  u64 A=pack2(a,a);
  for(int i=0;i<iters;i++){
    #pragma unroll
    for(int c=0;c<CH;c++) v[c]=sub2(A,v[c]);
  }
  float s=0; 
  #pragma unroll
  for(int c=0;c<CH;c++){ float x,y; unpack2(v[c],x,y); s+=x+y; }
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
// the 6-op distance chain + min3, register resident (no smem): upper bound for the matching loop
template<int S> __global__ void __launch_bounds__(256) peak_chain(float* out, int iters, float qa){
  float px[S],py[S],pz[S],m[S];
  #pragma unroll
  for(int s=0;s<S;s++){ px[s]=threadIdx.x*0.01f+s; py[s]=threadIdx.x*0.02f-s; pz[s]=s*0.5f; m[s]=1e5f; }
  u64 qx=pack2(qa,qa+1.f), qy=pack2(qa*2,qa*2+1.f), qz=pack2(qa*3,qa*3+1.f);
  u64 inc=pack2(0.25f,0.25f);
  for(int i=0;i<iters;i++){
    #pragma unroll
    for(int s=0;s<S;s++){
      u64 dx=sub2(pack2(px[s],px[s]),qx), dy=sub2(pack2(py[s],py[s]),qy), dz=sub2(pack2(pz[s],pz[s]),qz);
      u64 d=fma2(dz,dz,fma2(dx,dx,mul2(dy,dy)));
      float a,b; unpack2(d,a,b); m[s]=min3(m[s],a,b);
    }
    qx=sub2(qx,inc); qy=sub2(qy,inc); qz=sub2(qz,inc);
  }
  float s0=0;
  #pragma unroll
  for(int s=0;s<S;s++) s0+=m[s];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s0;
}
// scalar variant of the chain (FADD/FMUL/FFMA + FMNMX) for comparison
template<int S> __global__ void __launch_bounds__(256) peak_chain_scalar(float* out, int iters, float qa){
  float px[S],py[S],pz[S],m[S];
  #pragma unroll
  for(int s=0;s<S;s++){ px[s]=threadIdx.x*0.01f+s; py[s]=threadIdx.x*0.02f-s; pz[s]=s*0.5f; m[s]=1e5f; }
  float qx=qa, qy=qa*2, qz=qa*3;
  for(int i=0;i<iters;i++){
    #pragma unroll
    for(int s=0;s<S;s++){
      float dx=px[s]-qx, dy=py[s]-qy, dz=pz[s]-qz;
      float d=fmaf(dz,dz,fmaf(dx,dx,dy*dy));
      m[s]=fminf(m[s],d);
    }
    qx-=0.25f; qy-=0.25f; qz-=0.25f;
  }
  float s0=0;
  #pragma unroll
  for(int s=0;s<S;s++) s0+=m[s];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s0;
}

// ---- (2) matching loop prototype: S sources per thread, tile of TT targets in smem (SoA X|Y|Z)
template<int S, int THREADS, int TT>
__global__ void __launch_bounds__(THREADS) k1(const float* __restrict__ px, const float* __restrict__ py, const float* __restrict__ pz,
   const float4* __restrict__ tiles, int ntiles, float* __restrict__ out)
{
  extern __shared__ float4 sm[];
  float sx[S], sy[S], sz[S], m[S];
  int base = blockIdx.x*THREADS*S + threadIdx.x;
  #pragma unroll
  for (int s=0;s<S;s++){ sx[s]=px[base+s*THREADS]; sy[s]=py[base+s*THREADS]; sz[s]=pz[base+s*THREADS]; m[s]=1e5f; }
  for (int t=0;t<ntiles;t++){
    __syncthreads();
    for (int i=threadIdx.x;i<3*TT/4;i+=THREADS) sm[i]=tiles[(size_t)t*(3*TT/4)+i];
    __syncthreads();
    #pragma unroll 2
    for (int j=0;j<TT/4;j++){
      float4 X=sm[j], Y=sm[TT/4+j], Z=sm[2*TT/4+j];
      u64 x01=pack2(X.x,X.y), x23=pack2(X.z,X.w), y01=pack2(Y.x,Y.y), y23=pack2(Y.z,Y.w), z01=pack2(Z.x,Z.y), z23=pack2(Z.z,Z.w);
      #pragma unroll
      for (int s=0;s<S;s++){
        u64 PX=pack2(sx[s],sx[s]), PY=pack2(sy[s],sy[s]), PZ=pack2(sz[s],sz[s]);
        u64 dx=sub2(PX,x01), dy=sub2(PY,y01), dz=sub2(PZ,z01);
        u64 d=fma2(dz,dz,fma2(dx,dx,mul2(dy,dy)));
        float a,b; unpack2(d,a,b); m[s]=min3(m[s],a,b);
        dx=sub2(PX,x23); dy=sub2(PY,y23); dz=sub2(PZ,z23);
        d=fma2(dz,dz,fma2(dx,dx,mul2(dy,dy)));
        unpack2(d,a,b); m[s]=min3(m[s],a,b);
      }
    }
  }
  #pragma unroll
  for (int s=0;s<S;s++) out[base+s*THREADS]=m[s];
}

static float timeit(void(*launch)(void*), void* ctx, int reps){
  cudaEvent_t a,b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  launch(ctx); launch(ctx); CK(cudaDeviceSynchronize());
  float best=1e30f;
  for(int r=0;r<reps;r++){ CK(cudaEventRecord(a)); launch(ctx); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b)); float ms; CK(cudaEventElapsedTime(&ms,a,b)); if(ms<best)best=ms; }
  return best;
}
struct PeakCtx{ float* out; int iters; int kind; int blocks; };
template<int CH> void launch_peak(void* c){ PeakCtx* p=(PeakCtx*)c;
  if(p->kind==0) peak_ffma<CH><<<p->blocks,256>>>(p->out,p->iters,1.0001f,0.5f);
  else if(p->kind==1) peak_ffma2<CH><<<p->blocks,256>>>(p->out,p->iters,1.0001f,0.5f);
  else peak_fadd2<CH><<<p->blocks,256>>>(p->out,p->iters,0.7f); }
template<int S> void launch_chain(void* c){ PeakCtx* p=(PeakCtx*)c;
  if(p->kind==0) peak_chain<S><<<p->blocks,256>>>(p->out,p->iters,0.3f); else peak_chain_scalar<S><<<p->blocks,256>>>(p->out,p->iters,0.3f); }

struct K1Ctx{ float *px,*py,*pz,*out; float4* tiles; int N,M; };
template<int S,int THREADS,int TT> void launch_k1(void* c){ K1Ctx* p=(K1Ctx*)c;
  int blocks=p->N/(THREADS*S); size_t smem=3*TT*sizeof(float);
  cudaFuncSetAttribute(k1<S,THREADS,TT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k1<S,THREADS,TT><<<blocks,THREADS,smem>>>(p->px,p->py,p->pz,p->tiles,p->M/TT,p->out); }
template<int S,int THREADS,int TT> void run_k1(K1Ctx& c, double clk_ghz, int sms){
  float ms=timeit(launch_k1<S,THREADS,TT>,&c,5);
  double pairs=(double)c.N*c.M; double pps=pairs/(ms*1e-3);
  double peak=sms*128.0*2*clk_ghz*1e9;
  printf("k1 S=%2d thr=%4d TT=%5d blocks=%5d : %8.3f ms  %.3e pairs/s  %.2f TFLOP/s(8/pair)  %.1f%% of nominal %.1f TF\n",S,THREADS,TT,c.N/(THREADS*S),ms,pps,pps*8e-12,100*pps*8/peak,peak*1e-12);
}
int main(){
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop,0));
  int sms=prop.multiProcessorCount; int clk_khz=0; cudaDeviceGetAttribute(&clk_khz,cudaDevAttrClockRate,0); double clk=clk_khz*1e-6;
  printf("device %s SMs=%d clock=%.3f GHz\n",prop.name,sms,clk);
  float* out; CK(cudaMalloc(&out,sizeof(float)*sms*64*256));
  int iters=4096;
  for(int occ=1;occ<=8;occ*=2){
    PeakCtx pc{out,iters,0,sms*occ};
    for(int kind=0;kind<3;kind++){ pc.kind=kind;
      float ms=timeit(launch_peak<8>,&pc,5);
      double ops=(double)pc.blocks*256*8*iters*(kind==0?1:2);  // lane-ops
      double fl=ops*(kind==2?1:2);
      printf("peak %-6s CH=8 blocks/SM=%d: %7.3f ms  %.2f T lane-op/s  %.2f TFLOP/s  (%.2f lane-op/clk/SM @%.3fGHz)\n",kind==0?"FFMA":kind==1?"FFMA2":"FADD2",occ,ms,ops/(ms*1e-3)*1e-12,fl/(ms*1e-3)*1e-12,ops/(ms*1e-3)/(sms*clk*1e9),clk);
    }
  }
  for(int occ=1;occ<=4;occ*=2){
    PeakCtx pc{out,2048,0,sms*occ};
    for(int kind=0;kind<2;kind++){ pc.kind=kind;
      float ms=timeit(launch_chain<8>,&pc,5);
      double pairs=(double)pc.blocks*256*8*2048*(kind==0?2:1);
      printf("chain %-6s S=8 blocks/SM=%d: %7.3f ms  %.3e pairs/s  %.2f TFLOP/s(8/pair)\n",kind==0?"packed":"scalar",occ,ms,pairs/(ms*1e-3),pairs*8/(ms*1e-3)*1e-12);
      ms=timeit(launch_chain<4>,&pc,5);
      pairs=(double)pc.blocks*256*4*2048*(kind==0?2:1);
      printf("chain %-6s S=4 blocks/SM=%d: %7.3f ms  %.3e pairs/s  %.2f TFLOP/s(8/pair)\n",kind==0?"packed":"scalar",occ,ms,pairs/(ms*1e-3),pairs*8/(ms*1e-3)*1e-12);
    }
  }
  // matching prototype
  K1Ctx c; c.N=sms*4096; c.M=32768;
  std::vector<float> h(c.N); 
  CK(cudaMalloc(&c.px,4*c.N)); CK(cudaMalloc(&c.py,4*c.N)); CK(cudaMalloc(&c.pz,4*c.N)); CK(cudaMalloc(&c.out,4*c.N));
  for(int k=0;k<3;k++){ for(int i=0;i<c.N;i++) h[i]=(float)((i*2654435761u>>8)&0xffff)/16384.f-2.f+k; CK(cudaMemcpy(k==0?c.px:k==1?c.py:c.pz,h.data(),4*c.N,cudaMemcpyHostToDevice)); }
  std::vector<float> ht(3*c.M); for(size_t i=0;i<ht.size();i++) ht[i]=(float)((i*2246822519u>>8)&0xffff)/16384.f-2.f;
  CK(cudaMalloc(&c.tiles,4*3*c.M)); CK(cudaMemcpy(c.tiles,ht.data(),4*3*c.M,cudaMemcpyHostToDevice));
  run_k1<4,128,1024>(c,clk,sms); run_k1<4,256,1024>(c,clk,sms); run_k1<4,512,1024>(c,clk,sms);
  run_k1<8,128,1024>(c,clk,sms); run_k1<8,256,1024>(c,clk,sms); run_k1<8,512,1024>(c,clk,sms);
  run_k1<16,128,1024>(c,clk,sms); run_k1<16,256,1024>(c,clk,sms);
  run_k1<8,256,2048>(c,clk,sms); run_k1<8,256,4096>(c,clk,sms); run_k1<8,256,512>(c,clk,sms);
  run_k1<2,256,1024>(c,clk,sms); run_k1<2,512,1024>(c,clk,sms);
  printf("done\n");
  return 0;
}
