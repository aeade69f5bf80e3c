/* Test helper (tests/test_lowering.py): lol_smin_c as lol_kernel.cuh writes it -- the division by k as
 * n * rkh with two FMA corrections on n = b - a, rkh = RN(1/k)/2, k2 = 2k, and b - n*h for b + (a-b)*h --
 * against sminf (float.h:29-33) as the reference computes it, bit for bit, on random, nearly equal,
 * equal, signed-zero and tiny operands inside the guard's range. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
static float clamp01(float v){ v = v > 0.f ? v : 0.f; /* maxss(v,0): NaN -> 0 */ return v < 1.f ? v : 1.f; }
static float sat(float v){ return (v > 0.f) ? ((v < 1.f) ? v : 1.f) : 0.f; }
static float ref(float a, float b, float k){ float h = clamp01(.5f + (.5f*(b-a))/k); return (b + (a-b)*h) - (k*h)*(1.f-h); }
static float fast(float a, float b, float k, float rkh, float k2){ float n=b-a; float q0=n*rkh; float q=fmaf(fmaf(-k2,q0,n),rkh,q0); float h=sat(.5f+q); return (b - n*h) - (k*h)*(1.f-h); }
static uint64_t s=88172645463325252ull; static uint64_t rnd(){ s^=s<<13; s^=s>>7; s^=s<<17; return s; }
static float bits(uint32_t u){ float f; memcpy(&f,&u,4); return f; }
int main(int argc, char** argv){ long N = argc > 1 ? atol(argv[1]) : 3000000;
  float ks[] = {3.f, 4.f, 0.3f, 0.5f, 1.0f, 0.05f, 3.9f, 1e-5f, 1e5f, 0.7312f};
  long bad=0, n=0;
  for (int ki=0; ki<10; ki++){ float k=ks[ki], rkh=(1.0f/k)*0.5f, k2=k*2.0f;
    for (long i=0;i<N;i++){
      uint64_t r=rnd(); float a,b;
      int mode=i%6;
      if(mode==0){ a=bits((uint32_t)r); b=bits((uint32_t)(r>>32)); }
      else if(mode==1){ a=bits((uint32_t)r); b=a+ (float)((int)(r>>40)%2001-1000)*k*0.002f; }
      else if(mode==2){ a=(float)((double)(r&0xffffff)/1e5-80); b=(float)((double)((r>>24)&0xffffff)/1e5-80); }
      else if(mode==3){ a=bits((uint32_t)r & 0x807fffffu | ((uint32_t)(r>>33)%40)<<23); b=bits((uint32_t)(r>>32)&0x807fffffu | ((uint32_t)(r>>50)%40)<<23); }
      else if(mode==4){ a=bits((uint32_t)r); b=a; if(r>>63) b=-b; }
      else { a=(float)((double)(r&0xffff)/1e3); b=nextafterf(a, (r>>20)&1? 1e9f:-1e9f); }
      if(!(fabsf(a)<=0x1p61f && fabsf(b)<=0x1p61f)) continue;
      float x=ref(a,b,k), y=fast(a,b,k,rkh,k2); uint32_t ux,uy; memcpy(&ux,&x,4); memcpy(&uy,&y,4);
      n++;
      if(ux!=uy && !(x!=x && y!=y)){ if(bad<10) printf("k=%g a=%a b=%a ref=%a fast=%a\n",k,a,b,x,y); bad++; }
    }
  }
  printf("%ld cases, %ld mismatches\n", n, bad); return bad!=0;
}
