#include "../../vierkant_b200/csrc/host_copy.h"
#include <chrono>
#include <cstdio>
#include <cstdlib>
int main(int argc,char**argv){ size_t n=64u<<20; char*a=(char*)malloc(n),*b=(char*)aligned_alloc(4096,n); memset(a,1,n); memset(b,2,n);
 for(int w: {0,3,7}){ vkt::CopyPool p(w); for(int rep=0;rep<3;rep++){ auto t0=std::chrono::steady_clock::now(); for(size_t off=0;off<n;off+=4u<<20) p.copy(b+off,a+off,4u<<20); auto t1=std::chrono::steady_clock::now(); double ms=std::chrono::duration<double,std::milli>(t1-t0).count(); if(rep==2) printf("workers %d: %.2f ms %.1f GB/s ok=%d\n",w,ms,n/ms*1e-6,memcmp(a,b,n)==0);} } }
