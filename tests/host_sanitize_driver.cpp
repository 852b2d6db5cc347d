// Driver for tests/test_host_sanitizers.py: the host clustering / scoring entry points of libdd_b200 (leiden.cpp, louvain.cpp,
// score.cpp compiled as plain C++ with AddressSanitizer + UBSan) on kNN lists read from the working directory.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <string>
#include <vector>
#include "../include/dd_b200.h"
struct dd_handle;
void dd_set_global_error(const std::string &) {}
int dd_fail(dd_handle *, int code, const std::string &) { return code; }
template<class T> std::vector<T> rd(const char*f){FILE*fp=fopen(f,"rb");fseek(fp,0,SEEK_END);long s=ftell(fp);fseek(fp,0,SEEK_SET);std::vector<T> v(s/sizeof(T));if(fread(v.data(),1,s,fp)!=(size_t)s)abort();fclose(fp);return v;}
int main(){
  auto idx=rd<int32_t>("idx.bin"); auto dist=rd<float>("dist.bin"); auto idx31=rd<int32_t>("idx31.bin");
  const int64_t n=4000; const int k=10;
  std::vector<int32_t> lab(n); int32_t nc=0;
  int rc=dd_leiden_knn(n,k,idx.data(),dist.data(),4.0,0,lab.data(),&nc); printf("leiden rc=%d nc=%d\n",rc,nc);
  rc=dd_louvain_knn(n,k,idx.data(),4.0,0,lab.data(),&nc); printf("louvain_knn rc=%d nc=%d\n",rc,nc);
  rc=dd_phenograph_knn(n,31,idx31.data(),1,10,0,lab.data(),&nc); printf("phenograph rc=%d nc=%d\n",rc,nc);
  int64_t nnz=0; rc=dd_umap_connectivities(n,k,idx.data(),dist.data(),nullptr,nullptr,nullptr,0,&nnz); printf("umap rc=%d nnz=%lld\n",rc,(long long)nnz);
  std::vector<int64_t> ip(n+1); std::vector<int32_t> ix(nnz); std::vector<float> w(nnz);
  rc=dd_umap_connectivities(n,k,idx.data(),dist.data(),ip.data(),ix.data(),w.data(),nnz,&nnz);
  std::vector<int64_t> ix64(ix.begin(),ix.end()); std::vector<double> wd(w.begin(),w.end());
  rc=dd_louvain_csr_level0(n,ip.data(),ix64.data(),wd.data(),1.0,0,lab.data(),&nc); printf("louvain level0 weighted rc=%d nc=%d\n",rc,nc);
  rc=dd_louvain_csr_level0(n,ip.data(),ix64.data(),nullptr,4.0,0,lab.data(),&nc); printf("louvain level0 unit rc=%d nc=%d\n",rc,nc);
  rc=dd_leiden_csr(n,ip.data(),ix64.data(),wd.data(),1.0,1,lab.data(),&nc); printf("leiden_csr rc=%d nc=%d\n",rc,nc);
  rc=dd_louvain_csr(n,ip.data(),ix64.data(),wd.data(),1.0,1,lab.data(),&nc); printf("louvain_csr rc=%d nc=%d\n",rc,nc);
  std::vector<double> sc(3000), lp(3000); rc=dd_score(3000,1000,lab.data(),sc.data(),lp.data()); printf("score rc=%d\n",rc);
  // degenerate inputs
  int64_t ip0[2]={0,0}; rc=dd_leiden_csr(1,ip0,nullptr,nullptr,1.0,0,lab.data(),&nc); printf("leiden n=1 rc=%d nc=%d\n",rc,nc);
  int32_t i2[4]={0,1,1,0}; float d2[4]={0,0,0,0}; rc=dd_leiden_knn(2,2,i2,d2,1.0,0,lab.data(),&nc); printf("leiden n=2 zero dist rc=%d nc=%d\n",rc,nc);
  return 0; }
