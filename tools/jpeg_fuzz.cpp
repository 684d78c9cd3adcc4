// Mutation fuzz of the JPEG host decoder (fiducials_b200/csrc/jpeg_host.hpp) under the sanitizers:
//   cd tools && g++ -O1 -g -fsanitize=address,undefined -fno-sanitize-recover=all -o /tmp/jfuzz jpeg_fuzz.cpp && /tmp/jfuzz some.jpg
// 6000 mutated copies of a 1080p stream (byte flips, truncations, marker runs, damaged tables): clean (round 2).
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <random>
#include "../fiducials_b200/csrc/jpeg_host.hpp"
int main(int argc,char**argv){
  FILE*f=fopen(argv[1],"rb"); fseek(f,0,SEEK_END); long n=ftell(f); fseek(f,0,SEEK_SET); std::vector<uint8_t> base(n); if(fread(base.data(),1,n,f)!=(size_t)n) return 2; fclose(f);
  using namespace fidjpeg; std::mt19937 rng(123);
  // tight buffers so that ASAN sees any overrun: blocks for 1920x1088 4:2:0 (also the capacity passed in)
  size_t ok=0, bad=0;
  for(int it=0; it<6000; it++){
    std::vector<uint8_t> d=base;
    int mode=it%5;
    if(mode==0){ for(int k=0;k<1+(int)(rng()%6);k++) d[2+rng()%(d.size()-2)]=rng()&255; }
    else if(mode==1){ for(int k=0;k<1+(int)(rng()%30);k++) d[d.size()/3+rng()%(d.size()-d.size()/3)]=rng()&255; }
    else if(mode==2){ d.resize(4+rng()%(d.size()-4)); }
    else if(mode==3){ size_t p=2+rng()%(d.size()-10); for(int k=0;k<1+(int)(rng()%8);k++) d[p+k]=0xFF; }
    else { for(int k=0;k<1+(int)(rng()%4);k++) d[2+rng()%600]=rng()&255; }  // header area: tables, SOF, SOS
    FrameInfo fi; size_t mb=48960; std::vector<uint64_t> mask(mb); std::vector<uint32_t> off(mb); std::vector<int16_t> vals(mb*64); size_t nv=0;
    // exact-size heap copy of the input so that reads past the end are caught
    uint8_t* in=(uint8_t*)malloc(d.size()); memcpy(in,d.data(),d.size());
    int rc=decode_image(in,d.size(),&fi,mask.data(),off.data(),vals.data(),vals.size(),mb,&nv);
    free(in);
    if(rc==0) ok++; else bad++;
  }
  printf("ok %zu rejected %zu\n", ok, bad);
}
