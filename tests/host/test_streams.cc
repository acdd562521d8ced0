// tests/host/test_streams.cc -- exercises the drop-in C++ stream nodes (host/btk_streams.h) the way the reference's
// drivers use the original classes (btk/src/superdirectiveBeamformer.cc:150-220): build the chain, loop next() until
// jiterator_error, write what came out.  The Python tests compare the files with the oracle.
//
//   test_streams errors                      host-only checks (no GPU needed): exception types and codes
//   test_streams chain <in.bin> <out.bin>    DS (mode 0) or MVDR (mode 1) chain through the stream nodes;
//                                            mode 2: DS -> ZelinskiPostFilter -> synthesis (src/beamformerDS.cc:150-190);
//                                            mode 3: SubbandGSC with fixed active weights wa[s][k] = 0.05 (cos(s+k) + j sin(2s-k))
//                                            mode 4: null-steering DS (calcArrayManifoldVectors2, interferer delays = reversed
//                                                    target delays) -> frames pulled one by one, every frame's snapshots
//                                                    (getSnapShotArray) folded into a SpectralMatrixArray (mu 0.95), read
//                                                    after 10 frames and at the end; writes Y[4], R[M][C][C], w[B][C]
//   test_streams chain <in.bin> <out.bin> <file.wav>   mode 5: the channels come from ONE interleaved 16-bit WAVE file through
//                                            IterativeSampleFeature nodes (feature.cc:803-896), DS chain as in mode 0; the first
//                                            two blocks of channel 1 pulled in lock step are appended to the output
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../../distantspeechrecognition-mirror_b200/host/btk_streams.h"

using namespace btkb200;

static btk_vector make_vec(std::vector<double>& v) { btk_vector g; g.size = v.size(); g.stride = 1; g.data = v.data(); g.block = 0; g.owner = 0; return g; }

static int run_errors() {
  int fails = 0;
  std::vector<double> h(1024, 0.001);
  std::vector<float> x(1000, 1.f);
  btk_vector hv = make_vec(h);
  VectorFloatFeatureStreamPtr src(new MemorySampleFeature(x.data(), x.size(), 128, 128, true));
  try { OverSampledDFTAnalysisBank a(src, &hv, 256, 3, 1); fails++; }               // N = 768 != 1024
  catch (jconsistency_error& e) { if (e.getCode() != 3) fails++; }
  VectorFloatFeatureStreamPtr bad(new MemorySampleFeature(x.data(), x.size(), 100, 100, true));
  try { OverSampledDFTAnalysisBank a(bad, &hv, 256, 4, 1); fails++; }               // block length != D
  catch (jdimension_error& e) { if (e.getCode() != 4) fails++; }
  try { src->current(); fails++; } catch (jconsistency_error&) {}                   // frame index < 0
  int n = 0;
  try { for (;;) { src->next(); n++; } } catch (jiterator_error& e) { if (e.getCode() != 8) fails++; }   // literal: common/jexception.h:41-57, checked against the compiled reference in test_oracle_golden.py
  if (n != 8 || !src->isEnd()) fails++;                                             // ceil(1000/128) blocks, last one padded
  src->reset();
  if (src->frameX() != -1 || src->isEnd()) fails++;
  SubbandDS ds(256);
  std::vector<double> d(3, 0.0);
  btk_vector dv = make_vec(d);
  VectorComplexFeatureStreamPtr ch(new OverSampledDFTAnalysisBank(src, &hv, 256, 4, 1));
  ds.setChannel(ch);
  try { ds.calcArrayManifoldVectors(16000.0, &dv); fails++; } catch (jdimension_error&) {}   // 3 delays, 1 channel
  try { SubbandMVDR mv(512, true); fails++; } catch (j_error&) {}                   // halfBandShift unsupported (:2324-2327)
  printf("errors: %d failure(s)\n", fails);
  return fails;
}

int main(int argc, char** argv) {
  if (argc >= 2 && std::string(argv[1]) == "errors") return run_errors();
  if (argc < 4) { fprintf(stderr, "usage: test_streams errors | chain in.bin out.bin\n"); return 2; }
  FILE* f = fopen(argv[2], "rb");
  if (!f) return 2;
  int hdr[8];   // M m r dct C T mode reserved
  if (fread(hdr, sizeof(int), 8, f) != 8) return 2;
  const int M = hdr[0], m = hdr[1], r = hdr[2], dct = hdr[3], C = hdr[4], T = hdr[5], mode = hdr[6];
  const int N = M * m, D = M >> r;
  std::vector<double> h(N), g(N), tau(C), mic(C * 3);
  double load = 0, fs = 16000.0;
  std::vector<float> pcm((size_t)T * C);
  if (fread(h.data(), 8, N, f) != (size_t)N || fread(g.data(), 8, N, f) != (size_t)N || fread(tau.data(), 8, C, f) != (size_t)C ||
      fread(mic.data(), 8, C * 3, f) != (size_t)C * 3 || fread(&load, 8, 1, f) != 1 || fread(pcm.data(), 4, pcm.size(), f) != pcm.size()) return 2;
  fclose(f);
  try {
    btk_vector hv = make_vec(h), gv = make_vec(g), tv = make_vec(tau);
    std::shared_ptr<SubbandDS> bf(mode == 1 ? new SubbandMVDR(M) : mode == 3 ? new SubbandGSC(M) : new SubbandDS(M));
    std::vector<float> lockstep;
    if (mode == 5) {
      if (argc < 5) return 2;
      // lock-step pulling, the reference's use: every channel's node, block after block
      std::vector<std::shared_ptr<IterativeSampleFeature> > nodes;
      for (int c = 0; c < C; c++) nodes.push_back(std::shared_ptr<IterativeSampleFeature>(new IterativeSampleFeature(c, D, 0)));
      for (int c = 0; c < C; c++) { nodes[c]->reset(); nodes[c]->read(argv[4]); }
      for (int b = 0; b < 2; b++)
        for (int c = 0; c < C; c++) { const btk_vector_float* v = nodes[c]->next(); if (c == 1 % C) lockstep.insert(lockstep.end(), v->data, v->data + v->size); }
      if (nodes[0]->samplesN() == 0) return 3;
    }
    for (int c = 0; c < C; c++) {
      std::vector<float> x(T);
      for (int t = 0; t < T; t++) x[t] = pcm[(size_t)t * C + c];
      VectorFloatFeatureStreamPtr sample;
      if (mode == 5) {
        std::shared_ptr<IterativeSampleFeature> it(new IterativeSampleFeature(c, D, 0));
        it->reset();
        it->read(argv[4]);
        sample = it;
      } else {
        sample.reset(new MemorySampleFeature(x.data(), x.size(), D, D, true));
      }
      VectorComplexFeatureStreamPtr analysis(new OverSampledDFTAnalysisBank(sample, &hv, M, m, r, dct));
      bf->setChannel(analysis);
    }
    if (mode == 3) {
      SubbandGSC* gsc = static_cast<SubbandGSC*>(bf.get());
      gsc->calcGSCWeights(fs, &tv);
      std::vector<double> pw(2 * (C - 1));
      btk_vector pv = make_vec(pw);
      for (int s = 0; s <= M / 2; s++) {
        for (int k = 0; k < C - 1; k++) { pw[2 * k] = 0.05 * cos((double)(s + k)); pw[2 * k + 1] = 0.05 * sin((double)(2 * s - k)); }
        gsc->setActiveWeights_f(s, &pv);
      }
      if (gsc->getBlockingMatrix(0, 1)->size2 != (size_t)C - 1) return 3;
    } else if (mode != 4) {
      bf->calcArrayManifoldVectors(fs, &tv);
    }
    if (mode == 2 || mode == 3) {
      VectorComplexFeatureStreamPtr bfs = bf;
      VectorComplexFeatureStreamPtr last = bfs;
      ZelinskiPostFilterPtr pf;
      if (mode == 2) {
        pf.reset(new ZelinskiPostFilter(bfs, M, 0.6, TYPE_ZELINSKI1_ABS));
        pf->setBeamformer(bf);
        last = pf;
      }
      OverSampledDFTSynthesisBank synth(last, &gv, M, m, r, dct);
      std::vector<float> out;
      for (;;) {
        const btk_vector_float* b;
        try { b = synth.next(); } catch (jiterator_error&) { break; }
        out.insert(out.end(), b->data, b->data + b->size);
      }
      if (mode == 2 && (!pf->getPostFilterWeights() || pf->getPostFilterWeights()->size != (size_t)M)) return 3;
      FILE* o = fopen(argv[3], "wb");
      int oh[4] = {(int)out.size(), synth.fused() ? 1 : 0, 0, 0};
      fwrite(oh, sizeof(int), 4, o);
      fwrite(out.data(), 4, out.size(), o);
      fclose(o);
      printf("chain ok: %zu samples, mode=%d\n", out.size(), mode);
      return 0;
    }
    if (mode == 4) {
      std::vector<double> tj(tau.rbegin(), tau.rend());
      btk_vector jv = make_vec(tj);
      bf->calcArrayManifoldVectors2(fs, &tv, &jv);
      beamformerWeights* bw = bf->getBeamformerWeightObject(0);
      if (bw->fftLen() != (unsigned)M || bw->chanN() != (unsigned)C || bw->NC() != 2) return 3;
      SpectralMatrixArray sma(M, C, 0.95);
      sma.zero();
      std::vector<double> Y, R, W;
      int n = 0;
      for (;;) {
        const btk_vector_complex* y;
        try { y = bf->next(); } catch (jiterator_error&) { break; }
        if (n < 4) Y.insert(Y.end(), y->data, y->data + 2 * M);
        SnapShotArrayPtr sa = bf->getSnapShotArray();
        // the per-bin snapshots the node publishes agree with snapShotArray_f (beamformer.h:140-141)
        if (sa->getSnapShot(3)->data[0] != bf->snapShotArray_f(3)->data[0]) return 4;
        // feed the spectral matrix array channel by channel, as the reference's users do (newSample per channel, then update)
        std::vector<double> col(2 * (size_t)M);
        btk_vector_complex cv; cv.size = M; cv.stride = 1; cv.data = col.data(); cv.block = 0; cv.owner = 0;
        for (int c = 0; c < C; c++) {
          for (int s = 0; s < M; s++) { col[2 * s] = sa->getSnapShot(s)->data[2 * c]; col[2 * s + 1] = sa->getSnapShot(s)->data[2 * c + 1]; }
          sma.newSample(&cv, c);
        }
        sma.update();
        n++;
        if (n == 10) sma.getSpecMatrix(1);          // an intermediate read: the recursion continues from the folded state
      }
      for (int s = 0; s < M; s++) { const btk_matrix_complex* Rm = sma.getSpecMatrix(s); R.insert(R.end(), Rm->data, Rm->data + 2 * (size_t)C * C); }
      for (int s = 0; s <= M / 2; s++) { const btk_vector_complex* w = bf->getWeights(s); W.insert(W.end(), w->data, w->data + 2 * C); }
      FILE* o = fopen(argv[3], "wb");
      int oh[4] = {n, (int)Y.size(), (int)R.size(), (int)W.size()};
      fwrite(oh, sizeof(int), 4, o);
      fwrite(Y.data(), 8, Y.size(), o);
      fwrite(R.data(), 8, R.size(), o);
      fwrite(W.data(), 8, W.size(), o);
      fclose(o);
      printf("arrays ok: %d frames\n", n);
      return 0;
    }
    if (mode == 1) {
      SubbandMVDR* mv = static_cast<SubbandMVDR*>(bf.get());
      btk_matrix mp; mp.size1 = C; mp.size2 = 3; mp.tda = 3; mp.data = mic.data(); mp.block = 0; mp.owner = 0;
      if (!mv->setDiffuseNoiseModel(&mp, fs)) return 3;
      mv->setAllLevelsOfDiagonalLoading((float)load);
      if (!mv->calcMVDRWeights(fs, 1e-8)) return 3;
    }
    VectorComplexFeatureStreamPtr bfs = bf;
    OverSampledDFTSynthesisBank synth(bfs, &gv, M, m, r, dct);
    std::vector<float> out;
    for (;;) {
      const btk_vector_float* b;
      try { b = synth.next(); } catch (jiterator_error&) { break; }
      out.insert(out.end(), b->data, b->data + b->size);
    }
    const int fused = synth.fused() ? 1 : 0;
    // the beamformer node alone, after reset(): first 4 frames of Y (full M bins, complex double)
    synth.reset();
    std::vector<double> Y;
    for (int k = 0; k < 4; k++) { const btk_vector_complex* y = bf->next(); Y.insert(Y.end(), y->data, y->data + 2 * M); }
    // push-style synthesis (inputSourceVector + next) fed from the beamformer node: first 6 output frames
    bf->reset();
    OverSampledDFTSynthesisBank push(&gv, M, m, r, dct);
    std::vector<float> pushed;
    int fed = 0;
    while ((int)pushed.size() < 6 * D) {
      const btk_vector_complex* y;
      try { y = bf->next(); } catch (jiterator_error&) { break; }
      push.inputSourceVector(y); fed++;
      try { const btk_vector_float* b = push.next(); pushed.insert(pushed.end(), b->data, b->data + b->size); }
      catch (jiterator_error&) {}      // still priming: not enough frames yet
    }
    if (mode == 5) pushed = lockstep;
    FILE* o = fopen(argv[3], "wb");
    int oh[4] = {(int)out.size(), fused, (int)Y.size(), (int)pushed.size()};
    fwrite(oh, sizeof(int), 4, o);
    fwrite(out.data(), 4, out.size(), o);
    fwrite(Y.data(), 8, Y.size(), o);
    fwrite(pushed.data(), 4, pushed.size(), o);
    fclose(o);
    printf("chain ok: %zu samples, fused=%d, fed=%d\n", out.size(), fused, fed);
  } catch (std::exception& e) {
    fprintf(stderr, "exception: %s\n", e.what());
    return 1;
  }
  return 0;
}
