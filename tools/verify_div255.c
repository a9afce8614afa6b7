// exhaustive check: for all floats x in [0, 65536): fma-based division by 255 == IEEE x/255.0f
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <omp.h>
int main(){
  const float r = 1.0f/255.0f;
  uint32_t lo=0, hi; float top=65536.0f; memcpy(&hi,&top,4);
  long bad=0;
  #pragma omp parallel for reduction(+:bad) schedule(static)
  for (int64_t u=lo; u<(int64_t)hi; ++u){
    uint32_t b=(uint32_t)u; float x; memcpy(&x,&b,4);
    float q0 = x*r;
    float e = fmaf(-255.0f, q0, x);
    float q1 = fmaf(e, r, q0);
    float ref = x/255.0f;
    if (q1!=ref) bad++;
  }
  printf("bad=%ld of %u\n", bad, hi);
  return 0;
}
