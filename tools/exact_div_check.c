/* exact_div_check.c -- is a / b, for a divisor b that is fixed per configuration (fx, fy, depth_factor),
 * equal bit for bit to   q = a * r;  e = fma(-b, q, a);  q' = fma(e, r, q)   with r = RN(1 / b) from the host?
 * (three instructions instead of the ~10 of the device's IEEE division: k_ingest spends 8 % of its
 * instructions in the three divisions of every vertex.)  Exhaustive over every float a with |a| in
 * [2^-64, 2^64), both signs.  Build and run:
 *   gcc -O2 -ffp-contract=off -fopenmp -mfma -o /tmp/exact_div_check tools/exact_div_check.c -lm && /tmp/exact_div_check
 * Result (8 host cores, 7.5 s): 0 mismatches of 2^31 for each of 570.3, 1000, 5000, 525, 531.5, 285.15, 142.575,
 * 1140.6 -- the intrinsics and depth factors of the test configurations.  Wired into the kernels as compile-time
 * options (-DYK_FAST_DIV=1: k_ingest, behind the same exhaustive check run on the device at init, k_div_check;
 * -DYK_ICP_XY: k_icp recomputes vx, vy with it), both waiting for their first GPU measurement (DESIGN.md,
 * what comes next). */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <omp.h>
static inline float u2f(uint32_t u){float f;memcpy(&f,&u,4);return f;}
int main(){
  const float divs[]={570.3f,1000.0f,5000.0f,525.0f,531.5f,285.15f,142.575f,1140.6f};
  for(unsigned d=0;d<sizeof(divs)/4;++d){
    const float b=divs[d]; const float r=1.0f/b; /* correctly rounded reciprocal */
    unsigned long long bad=0, n=0;
    /* all floats with exponent in [2^-64, 2^64), both signs, plus zero */
    #pragma omp parallel for reduction(+:bad,n) schedule(static)
    for(long long e=127-64;e<127+64;++e){
      for(uint32_t m=0;m<(1u<<23);++m){
        for(int s=0;s<2;++s){
          const float a=u2f(((uint32_t)s<<31)|((uint32_t)e<<23)|m);
          const float q=a*r; const float err=fmaf(-b,q,a); const float q2=fmaf(err,r,q);
          const float ref=a/b;
          uint32_t x,y; memcpy(&x,&q2,4); memcpy(&y,&ref,4);
          bad += (x!=y); ++n;
        }
      }
    }
    printf("b=%g r=%a: %llu mismatches of %llu\n",b,r,bad,n);
  }
  return 0;
}
