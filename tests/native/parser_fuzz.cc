// Test-only driver: the host parser (all three modes: full parse, deferred tokens, deferred modes) on
// randomly corrupted frames, built with AddressSanitizer + UBSan by tests/test_parser_sanitizers.py.
// Every frame is parsed from an exact-size heap copy so that any over-read is reported.
//   parser_fuzz stream.ivf [trials per mode]
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <random>
#include <cstring>

#include "host/frame_parser.h"
namespace vp8r { void *HostAlloc(size_t b, bool){ return std::malloc(b);} void HostFree(void*p,bool){ std::free(p);} void DeviceFree(void*,int){} }
int main(int argc,char**argv){
  FILE*f=fopen(argv[1],"rb"); fseek(f,0,SEEK_END); long n=ftell(f); fseek(f,0,SEEK_SET); std::vector<uint8_t> d(n); if(fread(d.data(),1,n,f)!=(size_t)n) return 2; fclose(f);
  int trials = argc>2? atoi(argv[2]):200;
  std::vector<std::pair<size_t,size_t>> fr; size_t at=32; while(at+12<=d.size()){ uint32_t sz=d[at]|d[at+1]<<8|d[at+2]<<16|d[at+3]<<24; fr.push_back({at+12,sz}); at+=12+sz; }
  std::mt19937 rng(12345);
  long ok=0, bad=0;
  for(int mode=0; mode<3; ++mode)
  for(int t=0;t<trials;++t){
    vp8r::FrameParser p; p.set_defer_tokens(mode==1); p.set_defer_modes(mode==2);
    vp8r_frame out;
    for(auto&x:fr){
      std::vector<uint8_t> b(d.begin()+x.first, d.begin()+x.first+x.second);
      int m=rng()%4;
      if(m==0){ int k=1+rng()%8; for(int i=0;i<k;++i) b[rng()%b.size()]^=1u<<(rng()%8); }
      else if(m==1){ b.resize(1+rng()%b.size()); }
      else if(m==2){ size_t s=rng()%b.size(); for(size_t i=s;i<b.size()&&i<s+32;++i) b[i]=rng(); }
      // exact-size heap copy so that ASAN sees any over-read
      uint8_t *h=(uint8_t*)malloc(b.size()); memcpy(h,b.data(),b.size());
      int rc=p.Parse(h,b.size(),&out);
      free(h);
      if(rc) ++bad; else ++ok;
    }
  }
  printf("ok %ld rejected %ld\n", ok, bad);
}
