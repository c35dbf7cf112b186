// ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// C-ABI glue around the reference's own, unmodified Metadata<3> (rulebook
// builders).  The reference sources are compiled where they lie: this file
// #includes /root/reference/SparseConvNet/sparseconvnet/SCN/Metadata/Metadata.cpp
// through the -I path given by oracle/Makefile, with <google/dense_hash_map>
// resolved to oracle/shim/ (see that header for why).  Output goes to
// oracle/_ref/libscn_ref_rules.so only.
//
// The reference's pybind module (SCN/pybind.cpp:12-32) does not expose the
// rulebooks, but they are public members / public getters of Metadata
// (SCN/Metadata/Metadata.h:47-80,133-158), so parity tests read them here.
#include <torch/torch.h>

#include "Metadata/Metadata.cpp"
template class Metadata<3>;

namespace {
at::Tensor L3(const long *p) {
  auto t = torch::empty({3}, at::kLong);
  for (int i = 0; i < 3; i++)
    t.data_ptr<long>()[i] = p[i];
  return t;
}
} // namespace

extern "C" {

void *ref_md_create() { return new Metadata<3>(); }
void ref_md_destroy(void *m) { delete static_cast<Metadata<3> *>(m); }

// Metadata::inputLayer (Metadata.cpp:405-417).  Returns nActive at `spatial`.
long ref_md_input_layer(void *m_, const long *spatial, const long *coords, long nrows,
                        long ncols, long batchSize, long mode) {
  auto &m = *static_cast<Metadata<3> *>(m_);
  auto c = torch::from_blob(const_cast<long *>(coords), {nrows, ncols}, at::kLong).clone();
  auto sp = L3(spatial);
  m.inputLayer(sp, c, (Int)batchSize, (Int)mode);
  return m.getNActive(sp);
}
long ref_md_nactive(void *m_, const long *spatial) {
  return static_cast<Metadata<3> *>(m_)->getNActive(L3(spatial));
}
// Metadata::getSpatialLocations (Metadata.cpp:147-168); out is [nActive][4].
void ref_md_spatial_locations(void *m_, const long *spatial, long *out) {
  auto t = static_cast<Metadata<3> *>(m_)->getSpatialLocations(L3(spatial));
  std::memcpy(out, t.data_ptr<long>(), sizeof(long) * t.numel());
}
// Hash-iteration order of grid `spatial`, batch item b: ids (without ctr) in
// ascending bucket order.  Returns count.
long ref_md_iteration_order(void *m_, const long *spatial, long b, int *out) {
  auto &SGs = static_cast<Metadata<3> *>(m_)->getSparseGrid(L3(spatial));
  if (b >= (long)SGs.size())
    return 0;
  long n = 0;
  for (auto const &it : SGs[b].mp) {
    if (out)
      out[n] = it.second + SGs[b].ctr;
    n++;
  }
  return n;
}
long ref_md_batch_size(void *m_, const long *spatial) {
  return (long)static_cast<Metadata<3> *>(m_)->getSparseGrid(L3(spatial)).size();
}

// kind 0: inputLayerRuleBook; 1: getSubmanifoldRuleBook(a=spatial, b=filter);
// 2: getRuleBook(a=inSize, b=outSize, c=filter, d=stride); 3: getSparseToDenseRuleBook(a=spatial).  Returns RuleBook*.
void *ref_md_rulebook(void *m_, int kind, const long *a, const long *b, const long *c,
                      const long *d, int openmp) {
  auto &m = *static_cast<Metadata<3> *>(m_);
  if (kind == 0)
    return &m.inputLayerRuleBook;
  if (kind == 1)
    return &m.getSubmanifoldRuleBook(L3(a), L3(b), openmp != 0);
  if (kind == 3) // getSparseToDenseRuleBook(a = spatial size), Metadata.cpp:469-483
    return &m.getSparseToDenseRuleBook(L3(a), openmp != 0);
  return &m.getRuleBook(L3(a), L3(b), L3(c), L3(d), openmp != 0);
}
long ref_rb_nlists(void *rb) { return (long)static_cast<RuleBook *>(rb)->size(); }
long ref_rb_list_size(void *rb, long i) { return (long)(*static_cast<RuleBook *>(rb))[i].size(); }
void ref_rb_list_copy(void *rb, long i, int *dst) {
  auto &v = (*static_cast<RuleBook *>(rb))[i];
  std::memcpy(dst, v.data(), sizeof(int) * v.size());
}
// IntArrayHash<3> (Metadata/32bits.h:57-66), returned as the size_t it yields.
unsigned long ref_point_hash(int x, int y, int z) {
  Point<3> p = {x, y, z};
  return IntArrayHash<3>()(p);
}
}
