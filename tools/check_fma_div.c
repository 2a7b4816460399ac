#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <math.h>
#include <stdlib.h>
static inline float bf16f(uint16_t b){uint32_t u=((uint32_t)b)<<16;float f;memcpy(&f,&u,4);return f;}
static inline float halff(uint16_t h){ // fp16->fp32
  uint32_t s=(h>>15)&1,e=(h>>10)&31,m=h&1023;uint32_t u;
  if(e==0){ if(m==0){u=s<<31;} else { float f=ldexpf((float)m,-24); if(s) f=-f; return f;} }
  else if(e==31){u=(s<<31)|0x7f800000|(m<<13);} else {u=(s<<31)|((e+112)<<23)|(m<<13);} float f;memcpy(&f,&u,4);return f;}
int main(int argc,char**argv){
  int mode=atoi(argv[1]); // 0 bf16, 1 fp16
  long bad_q=0,bad_i=0,tot=0;
  int n = mode==0?0x7f80:0x7c00;
  for(int a=1;a<n;a++){
    float amax= mode==0?bf16f(a):halff(a);
    volatile float s=amax/127.0f; volatile float y=1.0f/s; if(!(s>=0x1p-100f && s<=0x1p100f)) continue;
    for(int b=0;b<=a;b++){
      float x= mode==0?bf16f(b):halff(b);
      volatile float ref=x/s;
      float q0=x*y; float r=fmaf(-q0,s,x); float q1=fmaf(r,y,q0);
      tot++;
      if(q1!=ref){bad_q++; if(bad_q<10) printf("q mismatch amax=%g x=%g ref=%.9g q1=%.9g q0=%.9g\n",amax,x,ref,q1,q0);}
      if(nearbyintf(q1)!=nearbyintf(ref)) {bad_i++; if(bad_i<10) printf("INT mismatch amax=%g x=%g ref=%.9g q1=%.9g\n",amax,x,ref,q1);}
      // magic rounding check
      float t=q1+12582912.0f; uint32_t u; memcpy(&u,&t,4); int8_t q8=(int8_t)(u&0xff);
      if((int)q8!=(int)nearbyintf(ref)) {bad_i++; if(bad_i<10) printf("MAGIC mismatch amax=%g x=%g ref=%.9g\n",amax,x,ref);}
    }
  }
  printf("mode %d total %ld bad_q %ld bad_int %ld\n",mode,tot,bad_q,bad_i);
}
