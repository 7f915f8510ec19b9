// Calibration probe (not product code): measures FP64 ceilings on the box.
//  - DMMA.8x8x4 and DFMA issue throughput (register-resident loops)
//  - cuBLAS DGEMM / ZGEMM (the "measured FP64 peak" roofline denominators)
//  - cuSOLVER zgesvd / zgesvdj wall time (the numbers the Jacobi SVD must beat)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_fp64 probe_fp64.cu -lcublas -lcusolver
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cublas_v2.h>
#include <cusolverDn.h>
#include <cuComplex.h>

#define CK(x) do{ cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)

template<int CH>
__global__ void dmma_loop(double* out, int iters, double seed){
  double a = seed + threadIdx.x, b = seed*0.5 + threadIdx.x;
  double c[CH][2];
  #pragma unroll
  for(int i=0;i<CH;i++){ c[i][0]=i; c[i][1]=-i; }
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int i=0;i<CH;i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]),"+d"(c[i][1]) : "d"(a),"d"(b));
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<CH;i++) s += c[i][0]+c[i][1];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int CH>
__global__ void dfma_loop(double* out, int iters, double seed){
  double a = seed + threadIdx.x*1e-9, b = seed*0.5;
  double c[CH];
  #pragma unroll
  for(int i=0;i<CH;i++) c[i]=i;
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int i=0;i<CH;i++) c[i] = fma(c[i], a, b);
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<CH;i++) s += c[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
// fragment layout check for m8n8k4: C = A(8x4,row) * B(4x8,col)
__global__ void layout_check(const double* A, const double* B, double* C){
  int lane=threadIdx.x, g=lane>>2, t=lane&3;
  double a=A[g*4+t];      // A[row=g][k=t], row-major 8x4
  double b=B[t*8+g];      // B[k=t][n=g],  row-major 4x8
  double c0=0,c1=0;
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0),"+d"(c1) : "d"(a),"d"(b));
  C[g*8+t*2]=c0; C[g*8+t*2+1]=c1;  // C[row=g][col=2t,2t+1]
}

int main(){
  int dev=0; CK(cudaSetDevice(dev));
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,dev));
  int clk=0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
  printf("{\"device\":\"%s\",\"sms\":%d,\"clock_khz\":%d}\n", p.name, p.multiProcessorCount, clk);
  cudaEvent_t e0,e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  double* out; CK(cudaMalloc(&out, sizeof(double)*148*16*1024));
  // layout check
  {
    double hA[32],hB[32],hC[64],ref[64]; for(int i=0;i<32;i++){hA[i]=1+i*0.37; hB[i]=2-i*0.11;}
    for(int r=0;r<8;r++)for(int c=0;c<8;c++){double s=0; for(int k=0;k<4;k++) s+=hA[r*4+k]*hB[k*8+c]; ref[r*8+c]=s;}
    double *dA,*dB,*dC; CK(cudaMalloc(&dA,256)); CK(cudaMalloc(&dB,256)); CK(cudaMalloc(&dC,512));
    CK(cudaMemcpy(dA,hA,256,cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB,hB,256,cudaMemcpyHostToDevice));
    layout_check<<<1,32>>>(dA,dB,dC); CK(cudaMemcpy(hC,dC,512,cudaMemcpyDeviceToHost));
    double err=0; for(int i=0;i<64;i++) err=fmax(err,fabs(hC[i]-ref[i]));
    printf("{\"probe\":\"dmma_layout_m8n8k4\",\"max_err\":%g}\n", err);
  }
  // DMMA throughput, vary warps/SM
  for(int warps : {4,8,16,32}){
    int iters=20000; const int CH=8;
    int blocks=p.multiProcessorCount, threads=warps*32;
    dmma_loop<CH><<<blocks,threads>>>(out,100,1.0); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0)); dmma_loop<CH><<<blocks,threads>>>(out,iters,1.0); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms,e0,e1);
    double flops = 2.0*256*CH*(double)iters*warps*blocks;
    printf("{\"probe\":\"dmma_884\",\"warps_per_sm\":%d,\"chains\":%d,\"ms\":%.3f,\"tflops\":%.2f}\n", warps, CH, ms, flops/ms/1e9);
  }
  for(int warps : {8,16,32}){
    int iters=20000; const int CH=16;
    int blocks=p.multiProcessorCount, threads=warps*32;
    dfma_loop<CH><<<blocks,threads>>>(out,100,1.0); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0)); dfma_loop<CH><<<blocks,threads>>>(out,iters,1.0000001); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms,e0,e1);
    double flops = 2.0*32*CH*(double)iters*warps*blocks;
    printf("{\"probe\":\"dfma\",\"warps_per_sm\":%d,\"chains\":%d,\"ms\":%.3f,\"tflops\":%.2f}\n", warps, CH, ms, flops/ms/1e9);
  }
  // cuBLAS
  cublasHandle_t h; cublasCreate(&h);
  for(int n : {2048, 4096, 8192}){
    size_t bytes=sizeof(cuDoubleComplex)*(size_t)n*n;
    cuDoubleComplex *A,*B,*C; CK(cudaMalloc(&A,bytes)); CK(cudaMalloc(&B,bytes)); CK(cudaMalloc(&C,bytes));
    CK(cudaMemset(A,0,bytes)); CK(cudaMemset(B,0,bytes));
    std::vector<double> hx(2*(size_t)n*n); for(size_t i=0;i<hx.size();i++) hx[i]=(double)rand()/RAND_MAX-0.5;
    CK(cudaMemcpy(A,hx.data(),bytes,cudaMemcpyHostToDevice)); CK(cudaMemcpy(B,hx.data(),bytes,cudaMemcpyHostToDevice));
    cuDoubleComplex one=make_cuDoubleComplex(1,0), zero=make_cuDoubleComplex(0,0);
    double done=1, dzero=0;
    float best=1e30f;
    for(int r=0;r<5;r++){
      CK(cudaEventRecord(e0)); cublasZgemm(h,CUBLAS_OP_N,CUBLAS_OP_N,n,n,n,&one,A,n,B,n,&zero,C,n); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; cudaEventElapsedTime(&ms,e0,e1); if(r>0 && ms<best) best=ms;
    }
    printf("{\"probe\":\"cublas_zgemm\",\"n\":%d,\"ms\":%.3f,\"tflops\":%.2f}\n", n, best, 8.0*n*(double)n*n/best/1e9);
    best=1e30f;
    for(int r=0;r<5;r++){
      CK(cudaEventRecord(e0)); cublasDgemm(h,CUBLAS_OP_N,CUBLAS_OP_N,n,n,2*n,&done,(double*)A,n,(double*)B,2*n,&dzero,(double*)C,n); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; cudaEventElapsedTime(&ms,e0,e1); if(r>0 && ms<best) best=ms;
    }
    printf("{\"probe\":\"cublas_dgemm\",\"m\":%d,\"n\":%d,\"k\":%d,\"ms\":%.3f,\"tflops\":%.2f}\n", n,n,2*n, best, 2.0*n*(double)n*2*n/best/1e9);
    if(n==8192){ // sustained 3 s
      CK(cudaEventRecord(e0)); int reps=0; float tot=0;
      while(tot<3000){ for(int r=0;r<3;r++) cublasZgemm(h,CUBLAS_OP_N,CUBLAS_OP_N,n,n,n,&one,A,n,B,n,&zero,C,n); reps+=3; CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&tot,e0,e1);}
      printf("{\"probe\":\"cublas_zgemm_sustained\",\"n\":%d,\"reps\":%d,\"ms_total\":%.1f,\"tflops\":%.2f}\n", n, reps, tot, 8.0*n*(double)n*n*reps/tot/1e9);
    }
    cudaFree(A); cudaFree(B); cudaFree(C);
  }
  // cuSOLVER SVD reference timings
  cusolverDnHandle_t sh; cusolverDnCreate(&sh);
  for(int n : {512, 1024, 2048}){
    size_t bytes=sizeof(cuDoubleComplex)*(size_t)n*n;
    cuDoubleComplex *A,*U,*V; double* S; int* info; CK(cudaMalloc(&A,bytes)); CK(cudaMalloc(&U,bytes)); CK(cudaMalloc(&V,bytes)); CK(cudaMalloc(&S,8*n)); CK(cudaMalloc(&info,4));
    std::vector<double> hx(2*(size_t)n*n); for(size_t i=0;i<hx.size();i++) hx[i]=(double)rand()/RAND_MAX-0.5;
    { int lwork=0; cusolverDnZgesvd_bufferSize(sh,n,n,&lwork); cuDoubleComplex* work; CK(cudaMalloc(&work,sizeof(cuDoubleComplex)*(size_t)lwork)); double* rwork; CK(cudaMalloc(&rwork,8*n));
      CK(cudaMemcpy(A,hx.data(),bytes,cudaMemcpyHostToDevice)); CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0)); cusolverDnZgesvd(sh,'S','S',n,n,A,n,S,U,n,V,n,work,lwork,rwork,info); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; cudaEventElapsedTime(&ms,e0,e1); printf("{\"probe\":\"cusolver_zgesvd\",\"n\":%d,\"ms\":%.2f}\n", n, ms); cudaFree(work); cudaFree(rwork); }
    { gesvdjInfo_t params; cusolverDnCreateGesvdjInfo(&params); cusolverDnXgesvdjSetTolerance(params,1e-14); cusolverDnXgesvdjSetMaxSweeps(params,30);
      int lwork=0; cusolverDnZgesvdj_bufferSize(sh,CUSOLVER_EIG_MODE_VECTOR,1,n,n,A,n,S,U,n,V,n,&lwork,params); cuDoubleComplex* work; CK(cudaMalloc(&work,sizeof(cuDoubleComplex)*(size_t)lwork));
      CK(cudaMemcpy(A,hx.data(),bytes,cudaMemcpyHostToDevice)); CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0)); cusolverDnZgesvdj(sh,CUSOLVER_EIG_MODE_VECTOR,1,n,n,A,n,S,U,n,V,n,work,lwork,info,params); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; cudaEventElapsedTime(&ms,e0,e1); int sweeps=0; cusolverDnXgesvdjGetSweeps(sh,params,&sweeps);
      printf("{\"probe\":\"cusolver_zgesvdj\",\"n\":%d,\"ms\":%.2f,\"sweeps\":%d}\n", n, ms, sweeps); cudaFree(work); }
    cudaFree(A); cudaFree(U); cudaFree(V); cudaFree(S); cudaFree(info);
  }
  return 0;
}
