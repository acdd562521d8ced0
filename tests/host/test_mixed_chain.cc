// tests/host/test_mixed_chain.cc -- the drop-in nodes compiled INSIDE the reference tree (BTKB200_WITH_BTK): B200 nodes
// and the reference's own nodes in one chain, wired like the shipped drivers (btk/src/beamformerDS.cc:150-190,
// btk/src/superdirectiveBeamformer.cc:150-205), against the all-reference chain on the same input.
//
// TEST INFRASTRUCTURE.  Built by oracle/Makefile (target _ref/mixed_chain) when /root/reference is present: it includes
// the reference's stream/stream.h, beamformer/beamformer.h, postfilter/postfilter.h and modulated/modulated.h, links the
// reference objects of oracle/_ref/libbtk_ref.so and the product's libbtkb200.so.  The binary travels to the GPU box.
//
//   mixed_chain errors                 no GPU needed: pointer / exception types are the reference's own
//   mixed_chain run <in.bin>           needs a GPU: prints one "name snr_db frames" line per mixed chain, exit 0 when
//                                      every chain is >= 70 dB against the all-reference chain
//
// Chains (B = B200 node, R = reference node):
//   zel_B_R_B   B banks -> B SubbandDS -> R ZelinskiPostFilter (postfilter.cc:340-500, reads the beamformer's SnapShotArray
//               and beamformerWeights through setBeamformer(SubbandDSPtr&)) -> B synthesis bank
//   ds_B_B_R    B banks -> B SubbandDS -> R OverSampledDFTSynthesisBank (modulated.cc:521-664)
//   ds_R_B_B    R OverSampledDFTAnalysisBank x C -> B SubbandDS -> B synthesis bank
//   ds_R_B_R    R banks -> B SubbandDS -> R synthesis bank
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <string>
#include <vector>

#define BTKB200_WITH_BTK 1
#include "../../distantspeechrecognition-mirror_b200/host/btk_streams.h"
#include "modulated/modulated.h"
#include "postfilter/postfilter.h"

namespace b2 = btkb200;

static gsl_vector* make_vector(const std::vector<double>& v) {
  gsl_vector* g = gsl_vector_alloc(v.size());
  for (size_t i = 0; i < v.size(); i++) gsl_vector_set(g, i, v[i]);
  return g;
}

struct Input {
  int M, m, r, dct, C, T;
  std::vector<double> h, g, tau;
  std::vector<float> pcm;
  std::vector<std::vector<float> > chan;
};

static bool load(const char* fn, Input& in) {
  FILE* f = fopen(fn, "rb");
  if (!f) return false;
  int hdr[8];
  if (fread(hdr, sizeof(int), 8, f) != 8) return false;
  in.M = hdr[0]; in.m = hdr[1]; in.r = hdr[2]; in.dct = hdr[3]; in.C = hdr[4]; in.T = hdr[5];
  const int N = in.M * in.m;
  in.h.resize(N); in.g.resize(N); in.tau.resize(in.C); in.pcm.resize((size_t)in.T * in.C);
  bool ok = fread(in.h.data(), 8, N, f) == (size_t)N && fread(in.g.data(), 8, N, f) == (size_t)N &&
            fread(in.tau.data(), 8, in.C, f) == (size_t)in.C && fread(in.pcm.data(), 4, in.pcm.size(), f) == in.pcm.size();
  fclose(f);
  in.chan.assign(in.C, std::vector<float>(in.T));
  for (int c = 0; c < in.C; c++) for (int t = 0; t < in.T; t++) in.chan[c][t] = in.pcm[(size_t)t * in.C + c];
  return ok;
}

template <class SynthPtr>
static std::vector<float> drain(SynthPtr& syn) {
  std::vector<float> out;
  for (;;) {
    const gsl_vector_float* b;
    try { b = syn->next(); } catch (jiterator_error&) { break; }
    for (size_t i = 0; i < b->size; i++) out.push_back(gsl_vector_float_get(b, i));
  }
  return out;
}

static double snr_db(const std::vector<float>& a, const std::vector<float>& ref) {
  if (a.size() != ref.size() || a.empty()) return -1000.0;
  double s = 0, e = 0;
  for (size_t i = 0; i < a.size(); i++) { s += (double)ref[i] * ref[i]; e += ((double)a[i] - ref[i]) * ((double)a[i] - ref[i]); }
  return e == 0 ? 300.0 : 10.0 * log10(s / e);
}

// kind of every stage: 'B' or 'R'; zel: put a ZelinskiPostFilter between beamformer and synthesis
static std::vector<float> run_chain(const Input& in, char banks, char bf_kind, char pf_kind, char syn_kind, bool* fused) {
  const int M = in.M, D = in.M >> in.r;
  gsl_vector* hv = make_vector(in.h);
  gsl_vector* gv = make_vector(in.g);
  gsl_vector* tv = make_vector(in.tau);
  // the beamformer is held through the REFERENCE's pointer type in every case
  ::SubbandDSPtr bf(bf_kind == 'B' ? static_cast< ::SubbandDS*>(new b2::SubbandDS(M)) : new ::SubbandDS(M, false));
  for (int c = 0; c < in.C; c++) {
    VectorFloatFeatureStreamPtr src(new b2::MemorySampleFeature(in.chan[c].data(), in.chan[c].size(), D, D, true));
    VectorComplexFeatureStreamPtr bank(banks == 'B'
        ? static_cast<VectorComplexFeatureStream*>(new b2::OverSampledDFTAnalysisBank(src, hv, M, in.m, in.r, in.dct))
        : static_cast<VectorComplexFeatureStream*>(new ::OverSampledDFTAnalysisBank(src, hv, M, in.m, in.r, in.dct)));
    bf->setChannel(bank);
  }
  bf->calcArrayManifoldVectors(16000.0, tv);          // virtual: the B200 node installs device weights AND the reference's object
  VectorComplexFeatureStreamPtr last((VectorComplexFeatureStreamPtr&)bf);
  ::ZelinskiPostFilterPtr pf;
  if (pf_kind == 'R') {
    pf = new ::ZelinskiPostFilter((VectorComplexFeatureStreamPtr&)bf, M, 0.6, (int)TYPE_ZELINSKI1_ABS, 0);
    pf->setBeamformer(bf);
    last = (VectorComplexFeatureStreamPtr&)pf;
  }
  std::vector<float> out;
  if (syn_kind == 'B') {
    b2::OverSampledDFTSynthesisBankPtr syn(new b2::OverSampledDFTSynthesisBank(last, gv, M, in.m, in.r, in.dct));
    out = drain(syn);
    if (fused) *fused = syn->fused();
  } else {
    ::OverSampledDFTSynthesisBankPtr syn(new ::OverSampledDFTSynthesisBank(last, gv, M, in.m, in.r, in.dct));
    out = drain(syn);
  }
  gsl_vector_free(hv); gsl_vector_free(gv); gsl_vector_free(tv);
  return out;
}

static int run_errors() {
  int fails = 0;
  std::vector<double> h(1024, 0.001);
  std::vector<float> x(1000, 1.f);
  gsl_vector* hv = make_vector(h);
  // a B200 source held by the reference's pointer type; end of stream is the reference's jiterator_error, code JITERATOR
  VectorFloatFeatureStreamPtr src(new b2::MemorySampleFeature(x.data(), x.size(), 128, 128, true));
  int n = 0;
  try { for (;;) { src->next(); n++; } } catch (::jiterator_error& e) { if (e.getCode() != JITERATOR) fails++; }
  if (n != 8 || !src->isEnd()) fails++;
  src->reset();
  // B200 analysis bank: constructor errors are the reference's classes
  try { b2::OverSampledDFTAnalysisBank a(src, hv, 256, 3, 1); fails++; } catch (::jconsistency_error& e) { if (e.getCode() != JCONSISTENCY) fails++; }
  // a B200 bank is a ::VectorComplexFeatureStream: the REFERENCE's SubbandDS accepts it as a channel
  VectorComplexFeatureStreamPtr bank(new b2::OverSampledDFTAnalysisBank(src, hv, 256, 4, 1));
  if (!bank.unique()) fails++;                                   // intrusive count of common/refcount.h:186-199
  {
    ::SubbandDS ref_ds(256, false);
    ref_ds.setChannel(bank);
    if (ref_ds.chanN() != 1 || bank.unique()) fails++;
  }
  // a B200 beamformer IS a ::SubbandDS: the reference's post-filter takes it
  ::SubbandDSPtr bf(new b2::SubbandDS(256));
  bf->setChannel(bank);
  gsl_vector* d3 = gsl_vector_calloc(3);
  try { bf->calcArrayManifoldVectors(16000.0, d3); fails++; } catch (::jdimension_error& e) { if (e.getCode() != JDIMENSION) fails++; }
  ::ZelinskiPostFilterPtr pf(new ::ZelinskiPostFilter((VectorComplexFeatureStreamPtr&)bf, 256, 0.6, (int)TYPE_ZELINSKI1_ABS, 0));
  pf->setBeamformer(bf);
  gsl_vector_free(d3); gsl_vector_free(hv);
  printf("errors: %d failure(s)\n", fails);
  return fails;
}

int main(int argc, char** argv) {
  if (argc >= 2 && std::string(argv[1]) == "errors") return run_errors();
  if (argc < 3 || std::string(argv[1]) != "run") { fprintf(stderr, "usage: mixed_chain errors | run in.bin\n"); return 2; }
  Input in;
  if (!load(argv[2], in)) { fprintf(stderr, "cannot read %s\n", argv[2]); return 2; }
  int bad = 0;
  try {
    const std::vector<float> ref_zel = run_chain(in, 'R', 'R', 'R', 'R', 0);
    const std::vector<float> ref_ds = run_chain(in, 'R', 'R', '-', 'R', 0);
    struct { const char* name; char banks, bf, pf, syn; const std::vector<float>* ref; } cases[] = {
        {"zel_B_R_B", 'B', 'B', 'R', 'B', &ref_zel}, {"ds_B_B_R", 'B', 'B', '-', 'R', &ref_ds},
        {"ds_R_B_B", 'R', 'B', '-', 'B', &ref_ds},   {"ds_R_B_R", 'R', 'B', '-', 'R', &ref_ds},
        {"ds_B_B_B", 'B', 'B', '-', 'B', &ref_ds}};
    for (size_t k = 0; k < sizeof(cases) / sizeof(cases[0]); k++) {
      bool fused = false;
      const std::vector<float> out = run_chain(in, cases[k].banks, cases[k].bf, cases[k].pf, cases[k].syn, &fused);
      const double s = snr_db(out, *cases[k].ref);
      printf("%s %.2f %zu fused=%d\n", cases[k].name, s, out.size() / (size_t)(in.M >> in.r), fused ? 1 : 0);
      if (!(s >= 70.0)) bad++;
    }
  } catch (std::exception& e) {
    fprintf(stderr, "exception: %s\n", e.what());
    return 1;
  }
  return bad;
}
