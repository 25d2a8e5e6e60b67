// host_driver.cpp — native test driver for the C++ host layer (rdis_b200/host/rdis_host.h).
// Reads a problem written by tests/test_host_adapter.py (flat little-endian arrays), builds the
// reference-shaped object graph (Variable / Factor / OptimizableFunction), and runs one of:
//   children <file>            ComponentBatcher::createChildren over the unassigned variables (no GPU needed)
//   children_gpu <file>        ComponentBatcher::createChildrenOnDevice: the same listing through rdisgpu_components
//   wave <file> <maxiters>     one sibling wave through CudaSubspaceOptimizer::optimizeBatch
//   benchwaves <bal> <steps> <warmup>   bench.py's end-to-end leg (timed alternating waves on a BAL file)
//   siblings[_seq] <file> <maxiters> <useBounds> <parentFmin> <childFmin0>   the sibling loop with branch & bound, batched / sequential
//   single <file> <maxiters>   the same wave, one CudaSubspaceOptimizer::optimize call per child
//                              (the reference's sibling loop, src/RDISOptimizer.cpp:291-314)
// Output is plain text with %.17g numbers; the Python side compares it with the ctypes path and the oracle.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <limits>
#include <memory>

#include "rdis_builders.h"
#include "rdis_host.h"
#include "../../include/rdis_gpu.h"

using namespace rdis;

namespace {
template <class T>
std::vector<T> rd(std::ifstream& in, size_t n) {
  std::vector<T> v(n);
  in.read(reinterpret_cast<char*>(v.data()), (std::streamsize)(n * sizeof(T)));
  if (!in) throw std::runtime_error("short read");
  return v;
}

struct Loaded {
  std::unique_ptr<OptimizableFunction> fn;
  std::vector<double> x0;
  std::vector<uint8_t> assigned;
};

Loaded load(const char* path) {
  std::ifstream in(path, std::ios::binary);
  if (!in) throw std::runtime_error("cannot open problem file");
  auto hdr = rd<int64_t>(in, 6);
  const int64_t kind = hdr[0], V = hdr[1], F = hdr[2], E = hdr[3], ncams = hdr[4], npts = hdr[5];
  auto lb = rd<double>(in, (size_t)V), ub = rd<double>(in, (size_t)V);
  Loaded L;
  L.fn.reset(new OptimizableFunction());
  for (int64_t i = 0; i < V; ++i) L.fn->addVariable(lb[(size_t)i], ub[(size_t)i]);
  if (kind == 0) {
    auto rowptr = rd<int64_t>(in, (size_t)F + 1);
    auto vid = rd<int32_t>(in, (size_t)E);
    auto expo = rd<double>(in, (size_t)E), konst = rd<double>(in, (size_t)E);
    auto sine = rd<uint8_t>(in, (size_t)E);
    auto coeff = rd<double>(in, (size_t)F);
    for (int64_t j = 0; j < F; ++j) {
      NonlinearProductFactor* f = L.fn->addProductFactor(coeff[(size_t)j]);
      for (int64_t e = rowptr[(size_t)j]; e < rowptr[(size_t)j + 1]; ++e)
        f->addVariable(L.fn->getVariables()[(size_t)vid[(size_t)e]], expo[(size_t)e], konst[(size_t)e], sine[(size_t)e] != 0);
    }
  } else {
    auto cam = rd<int32_t>(in, (size_t)F), pt = rd<int32_t>(in, (size_t)F);
    auto obs = rd<double>(in, 2 * (size_t)F);
    L.fn->declareBundleAdjustment((int32_t)ncams, (int32_t)npts);
    for (int64_t j = 0; j < F; ++j) L.fn->addObservation(cam[(size_t)j], pt[(size_t)j], obs[2 * (size_t)j], obs[2 * (size_t)j + 1]);
  }
  L.x0 = rd<double>(in, (size_t)V);
  L.assigned = rd<uint8_t>(in, (size_t)V);
  return L;
}

void assign_flagged(Loaded& L) {
  for (size_t i = 0; i < L.assigned.size(); ++i)
    if (L.assigned[i]) L.fn->getVariables()[i]->assign(L.x0[i]);
}

VariableIDVec unassigned(Loaded& L) {
  VariableIDVec v;
  for (Variable* var : L.fn->getVariables())
    if (!var->isAssigned()) v.push_back(var->getID());
  return v;
}
}  // namespace

int main(int argc, char** argv) {
  if (argc < 3) {
    std::fprintf(stderr, "usage: host_driver children|wave|single <file> [maxiters] [lm]\n");
    return 2;
  }
  try {
    if (std::string(argv[1]) == "loadbal") {  // loadbal <bal file> [ncams npts]: the BAL loader's view of the file
      BundleAdjustmentFunction fn;
      const long long nc = argc > 3 ? std::atoll(argv[3]) : -1, np = argc > 4 ? std::atoll(argv[4]) : -1;
      if (!fn.load(argv[2], nc, np)) {
        std::printf("load failed\n");
        return 1;
      }
      std::printf("V %lld F %zu ncams %lld npts %lld blocks %lld\n", fn.getNumVars(), fn.getFactors().size(), fn.getNumCameras(),
                  fn.getNumPoints(), fn.getNumBlocks());
      for (const Factor* f : fn.getFactors()) {
        const BundleAdjustmentFactor* b = static_cast<const BundleAdjustmentFactor*>(f);
        std::printf("f %d %d %.17g %.17g %lld %lld\n", b->getCamera(), b->getPoint(), b->obsX(), b->obsY(),
                    b->getVariables()[0]->getID(), b->getVariables()[9]->getID());
      }
      for (Variable* v : fn.getVariables())
        std::printf("v %.17g %.17g %.17g %.17g %.17g %lld\n", fn.getInitialState()[(size_t)v->getID()], v->getDomain().min(),
                    v->getDomain().max(), v->getDomain().samplingMin(), v->getDomain().samplingMax(), fn.getBlockID(v->getID()));
      return 0;
    }
    if (std::string(argv[1]) == "bawaves") {  // bawaves <bal file> <rounds> [lm]: alternating point / camera waves from the file state
      BundleAdjustmentFunction fn;
      if (!fn.load(argv[2])) return 1;
      const int rounds = argc > 3 ? std::atoi(argv[3]) : 2;
      const bool lm = argc > 4 && std::string(argv[4]) == "lm";
      fn.init(0);
      CudaSubspaceOptimizer cgd(fn);
      CudaLMSubspaceOptimizer lmo(fn);
      CudaSubspaceOptimizer& ssopt = lm ? static_cast<CudaSubspaceOptimizer&>(lmo) : cgd;
      ParameterMap opts;
      opts["SSmaxit"] = 25;  // the optBA default (src/RDISOptimizer.cpp:107)
      ssopt.setParameters(opts);
      const NumericVec& x0 = fn.getInitialState();
      VariablePtrVec& vars = fn.getVariables();
      for (Variable* v : vars) v->assign(x0[(size_t)v->getID()]);
      std::printf("objective %.17g\n", fn.eval());
      const VariableID npv = 9 * fn.getNumCameras();
      for (int rd = 0; rd < 2 * rounds; ++rd) {
        // un-assign one side (points on even half-rounds, cameras on odd): its blocks become the sibling components
        const bool points = (rd % 2 == 0);
        NumericVec cur((size_t)fn.getNumVars());
        for (Variable* v : vars) cur[(size_t)v->getID()] = v->eval();
        VariableIDVec open;
        for (Variable* v : vars)
          if ((v->getID() >= npv) == points) {
            v->unassign();
            open.push_back(v->getID());
          }
        std::vector<ChildComponent> kids;
        ComponentBatcher::createChildren(fn, open, kids);
        std::vector<ComponentProblem> probs(kids.size());
        for (size_t k = 0; k < kids.size(); ++k) ComponentBatcher::leafProblem(fn, kids[k], cur, probs[k]);
        const double total = ssopt.optimizeBatch(probs, false);
        std::printf("wave %d %s components %zu sum %.17g objective %.17g\n", rd, points ? "points" : "cameras", kids.size(), total,
                    fn.eval());
      }
      return 0;
    }
    if (std::string(argv[1]) == "benchwaves") {
      // benchwaves <bal file> <steps> <warmup> [strict]: bench.py's end-to-end leg.  One step = what the tree search
      // does for one alternating wave through the plugin surface with HOST objects: Variable::assign of the start state
      // (queued, uploaded by the adapter), CudaSubspaceOptimizer::optimizeBatch over the point components (index
      // lists + start values up, results down, Variable write-back), the same over the camera components.  Wall
      // clock around the steps; the sibling sets are built once before (their cost is printed as dispatch_ms).
      BundleAdjustmentFunction fn;
      if (!fn.load(argv[2])) return 1;
      const int steps = argc > 3 ? std::atoi(argv[3]) : 10, warmup = argc > 4 ? std::atoi(argv[4]) : 3;
      fn.init(0);
      if (argc > 5 && std::string(argv[5]) == "strict") rdisgpu_set_option(fn.device(), "strict", 1);
      CudaSubspaceOptimizer ssopt(fn);
      ParameterMap opts;
      opts["SSmaxit"] = 25;
      opts["SSftol"] = 3e-8;
      ssopt.setParameters(opts);
      const NumericVec& x0 = fn.getInitialState();
      VariablePtrVec& vars = fn.getVariables();
      for (Variable* v : vars) v->assign(x0[(size_t)v->getID()]);
      const VariableID npv = 9 * fn.getNumCameras();
      std::vector<ComponentProblem> waves[2];
      const auto d0 = std::chrono::steady_clock::now();
      for (int side = 0; side < 2; ++side) {  // 0: points open (cameras fixed), 1: cameras open
        VariableIDVec open;
        for (Variable* v : vars)
          if ((v->getID() >= npv) == (side == 0)) {
            v->unassign();
            open.push_back(v->getID());
          }
        std::vector<ChildComponent> kids;
        ComponentBatcher::createChildren(fn, open, kids);
        waves[side].resize(kids.size());
        for (size_t k = 0; k < kids.size(); ++k) ComponentBatcher::leafProblem(fn, kids[k], x0, waves[side][k]);
        for (VariableID vid : open) vars[(size_t)vid]->assign(x0[(size_t)vid]);
      }
      const double dispatch_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - d0).count();
      double total_ms = 0, objective = 0, pts_sum = 0, objective_first = 0;
      for (int it = 0; it < warmup + steps; ++it) {
        const auto t0 = std::chrono::steady_clock::now();
        for (Variable* v : vars) v->assign(x0[(size_t)v->getID()]);
        for (int side = 0; side < 2; ++side) {
          for (ComponentProblem& p : waves[side])
            for (size_t i = 0; i < p.vars.size(); ++i) p.xval[i] = (side == 0) ? x0[(size_t)p.vars[i]->getID()] : p.vars[i]->eval();
          const double tot = ssopt.optimizeBatch(waves[side], false);
          if (side == 0) pts_sum = tot; else objective = tot;
        }
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (it == 0) objective_first = objective;  // first visit: the batches are built and cached; later steps hit the cache
        if (it >= warmup) total_ms += ms;
        if (it + 1 == warmup) ssopt.resetTiming();
      }
      const CudaSubspaceOptimizer::Timing& tm = ssopt.timing();
      size_t nsolves = waves[0].size() + waves[1].size(), nv = 0, nf = 0;
      for (int side = 0; side < 2; ++side)
        for (const ComponentProblem& p : waves[side]) { nv += p.vars.size(); nf += p.factors.size(); }
      std::printf("{\"solves_per_step\": %zu, \"steps\": %d, \"warmup\": %d, \"ms_per_step\": %.6f, \"solves_per_s\": %.3f, "
                  "\"objective_after_step\": %.17g, \"objective_first_step\": %.17g, \"point_wave_sum\": %.17g, \"dispatch_ms\": %.3f, \"vars_in_problems\": %zu, "
                  "\"factors_in_problems\": %zu, \"V\": %lld, \"plugin_ms_per_step\": {\"recognise_and_pack\": %.4f, \"upload_assigned\": %.4f, \"device_calls\": %.4f, "
                  "\"of_which_fetch_wait\": %.4f, \"write_back\": %.4f}}\n",
                  nsolves, steps, warmup, total_ms / steps, nsolves * steps / (total_ms * 1e-3), objective, objective_first, pts_sum, dispatch_ms, nv, nf,
                  fn.getNumVars(), tm.pack_ms / steps, tm.flush_ms / steps, tm.device_ms / steps, tm.fetch_ms / steps, tm.writeback_ms / steps);
      return 0;
    }
    if (std::string(argv[1]) == "sinusoid_flat") {
      // sinusoid_flat <height> <branches> <maxArity> <odd>: builds the function, exports the flat arrays rdisgpu_add_nlpf
      // takes, prints their sizes, the build + export wall time and an FNV-1a hash of every array (no GPU needed):
      // the Python generator's arrays must hash the same
      if (argc < 6) return 2;
      const auto t0 = std::chrono::steady_clock::now();
      std::unique_ptr<OptimizableFunction> fn(makeHighDimSinusoid(std::atoll(argv[2]), std::atoll(argv[3]), std::atoll(argv[4]),
                                                                  std::atoi(argv[5]) != 0));
      const double build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
      std::vector<int64_t> rowptr;
      std::vector<int32_t> vid;
      std::vector<double> expo, konst, coeff;
      std::vector<uint8_t> sine;
      fn->exportProductFactors(rowptr, vid, expo, konst, sine, coeff);
      const double total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
      auto fnv = [](const void* p, size_t n) {
        unsigned long long h = 1469598103934665603ULL;
        const unsigned char* b = static_cast<const unsigned char*>(p);
        for (size_t i = 0; i < n; ++i) {
          h ^= b[i];
          h *= 1099511628211ULL;
        }
        return h;
      };
      std::printf("V %lld F %zu E %zu build_ms %.1f total_ms %.1f\n", fn->getNumVars(), coeff.size(), vid.size(), build_ms, total_ms);
      std::printf("hash rowptr %llx vid %llx expo %llx konst %llx sine %llx coeff %llx\n", fnv(rowptr.data(), rowptr.size() * 8),
                  fnv(vid.data(), vid.size() * 4), fnv(expo.data(), expo.size() * 8), fnv(konst.data(), konst.size() * 8),
                  fnv(sine.data(), sine.size()), fnv(coeff.data(), coeff.size() * 8));
      if (argc > 6) {  // ... <outfile>: the arrays themselves, little endian, in the order of the hash line
        std::ofstream out(argv[6], std::ios::binary);
        out.write(reinterpret_cast<const char*>(rowptr.data()), (std::streamsize)(rowptr.size() * 8));
        out.write(reinterpret_cast<const char*>(vid.data()), (std::streamsize)(vid.size() * 4));
        out.write(reinterpret_cast<const char*>(expo.data()), (std::streamsize)(expo.size() * 8));
        out.write(reinterpret_cast<const char*>(konst.data()), (std::streamsize)(konst.size() * 8));
        out.write(reinterpret_cast<const char*>(sine.data()), (std::streamsize)sine.size());
        out.write(reinterpret_cast<const char*>(coeff.data()), (std::streamsize)(coeff.size() * 8));
      }
      return 0;
    }
    if (std::string(argv[1]) == "sinusoid") {  // sinusoid <height> <branches> <maxArity> <odd>
      if (argc < 6) return 2;
      std::unique_ptr<OptimizableFunction> fn(makeHighDimSinusoid(std::atoll(argv[2]), std::atoll(argv[3]), std::atoll(argv[4]),
                                                                  std::atoi(argv[5]) != 0));
      std::printf("V %lld F %zu dom %.17g %.17g samp %.17g %.17g\n", fn->getNumVars(), fn->getFactors().size(),
                  fn->getVariables()[0]->getDomain().min(), fn->getVariables()[0]->getDomain().max(),
                  fn->getVariables()[0]->getDomain().samplingMin(), fn->getVariables()[0]->getDomain().samplingMax());
      for (const Factor* f : fn->getFactors()) {
        const NonlinearProductFactor* n = static_cast<const NonlinearProductFactor*>(f);
        std::printf("f %.17g %zu", n->getCoefficient(), n->numVars());
        for (size_t i = 0; i < n->numVars(); ++i)
          std::printf(" %lld %.17g %.17g %d", n->getVariables()[i]->getID(), n->getTerms()[i].exponent, n->getTerms()[i].constant,
                      n->getTerms()[i].useSine ? 1 : 0);
        std::printf("\n");
      }
      return 0;
    }
    Loaded L = load(argv[2]);
    const std::string mode = argv[1];
    assign_flagged(L);
    std::vector<ChildComponent> kids;
    if (mode == "children_gpu") {  // children_gpu <file>: the same listing, labelled by rdisgpu_components
      L.fn->init(0);
      ComponentBatcher::createChildrenOnDevice(*L.fn, unassigned(L), kids);
    } else {
      ComponentBatcher::createChildren(*L.fn, unassigned(L), kids);
    }
    if (mode == "children" || mode == "children_gpu") {
      std::printf("children %zu\n", kids.size());
      for (const ChildComponent& c : kids) {
        std::printf("%zu %zu |", c.vars.size(), c.factors.size());
        for (VariableID v : c.vars) std::printf(" %lld", v);
        std::printf(" |");
        for (FactorID f : c.factors) std::printf(" %lld", f);
        std::printf("\n");
      }
      return 0;
    }
    const int maxiters = argc > 3 ? std::atoi(argv[3]) : 25;
    L.fn->init(0);
    if (mode == "siblings" || mode == "siblings_seq") {
      // siblings[_seq] <file> <maxiters> <useBounds> <parentFmin> <childFmin0>: the sibling loop with branch & bound —
      // batched (ComponentBatcher::optimizeSiblings: one device call + replay) or the reference's own sequential form
      // (one CudaSubspaceOptimizer::optimize per child it actually visits, src/RDISOptimizer.cpp:291-314)
      const bool useBounds = argc > 4 && std::atoi(argv[4]) != 0;
      const double parentFmin = argc > 5 ? std::atof(argv[5]) : 0.0, childFmin0 = argc > 6 ? std::atof(argv[6]) : 0.0;
      CudaSubspaceOptimizer ssopt(*L.fn);
      ParameterMap opts;
      opts["SSmaxit"] = maxiters;
      opts["SSftol"] = 3e-8;
      ssopt.setParameters(opts);
      std::vector<ComponentProblem> probs(kids.size());
      std::vector<FactorPtrVec> lists(kids.size());
      for (size_t k = 0; k < kids.size(); ++k) {
        ComponentBatcher::leafProblem(*L.fn, kids[k], L.x0, probs[k]);
        for (FactorID f : kids[k].factors) lists[k].push_back(L.fn->getFactors()[(size_t)f]);
      }
      std::vector<NumericInterval> uab;
      L.fn->computeBoundsBatch(lists, uab);  // children's variables are unassigned here: Component::computeBounds at creation
      ComponentBatcher::SiblingWave w;
      if (mode == "siblings") {
        ComponentBatcher::optimizeSiblings(ssopt, *L.fn, probs, uab, parentFmin, childFmin0, useBounds, w);
      } else {
        const size_t n = probs.size();
        w.outcome.assign(n, ComponentBatcher::SIB_PRUNED);
        w.value.assign(n, std::numeric_limits<double>::quiet_NaN());
        double childFmin = childFmin0, assignedLB = 0;
        for (size_t k = 0; k < n; ++k) assignedLB += uab[k].lower();
        for (size_t k = 0; k < n; ++k) {
          const double fmin_k = childFmin + uab[k].lower();
          double fx;
          if (useBounds && fmin_k < uab[k].lower()) {
            w.outcome[k] = ComponentBatcher::SIB_BOUND_SKIPPED;
            fx = uab[k].lower();
          } else {
            w.outcome[k] = ComponentBatcher::SIB_OPTIMISED;
            fx = probs[k].fval = ssopt.optimize(probs[k].vars, probs[k].factors, probs[k].xval, probs[k].deltaFval, false);
          }
          w.value[k] = fx;
          ++w.evaluated;
          childFmin += uab[k].lower();
          childFmin -= fx;
          assignedLB += fx - uab[k].lower();
          if (k + 1 < n && useBounds && parentFmin <= assignedLB) break;
        }
        w.childFmin = childFmin;
        w.assignedLower = assignedLB;
      }
      std::printf("siblings %zu evaluated %zu childFmin %.17g assignedLower %.17g\n", probs.size(), w.evaluated, w.childFmin, w.assignedLower);
      for (size_t k = 0; k < probs.size(); ++k)
        std::printf("child %zu outcome %d value %.17g uab %.17g %.17g\n", k, w.outcome[k], w.value[k], uab[k].lower(), uab[k].upper());
      // the state the loop leaves behind: host objects and the device mirror (read back through the C-ABI)
      std::vector<double> dev((size_t)L.fn->getNumVars());
      rdisgpu_get_x(L.fn->device(), (int64_t)dev.size(), nullptr, dev.data());
      for (Variable* v : L.fn->getVariables()) {
        if (v->isAssigned()) std::printf("v %lld 1 %.17g %.17g\n", v->getID(), v->eval(), dev[(size_t)v->getID()]);
        else std::printf("v %lld 0\n", v->getID());
      }
      return 0;
    }
    const bool lm = argc > 4 && std::string(argv[4]) == "lm";
    CudaSubspaceOptimizer cgd(*L.fn);
    CudaLMSubspaceOptimizer lmo(*L.fn);
    CudaSubspaceOptimizer& ssopt = lm ? static_cast<CudaSubspaceOptimizer&>(lmo) : cgd;
    ParameterMap opts;
    opts["SSmaxit"] = maxiters;
    opts["SSftol"] = 3e-8;
    ssopt.setParameters(opts);
    std::vector<ComponentProblem> probs(kids.size());
    for (size_t k = 0; k < kids.size(); ++k) ComponentBatcher::leafProblem(*L.fn, kids[k], L.x0, probs[k]);
    double total = 0;
    if (mode == "wave") {
      total = ssopt.optimizeBatch(probs, false);
    } else if (mode == "single") {
      for (ComponentProblem& p : probs) {
        p.fval = ssopt.optimize(p.vars, p.factors, p.xval, p.deltaFval, false);
        total += p.fval;
        // the reference un-assigns the quick-assigned values before the next sibling (quickUnassignSSInitialVal,
        // src/RDISOptimizer.cpp:1184); siblings share no factor, so leaving them assigned changes nothing
      }
    } else {
      std::fprintf(stderr, "unknown mode\n");
      return 2;
    }
    std::printf("problems %zu total %.17g\n", probs.size(), total);
    for (const ComponentProblem& p : probs) {
      std::printf("%.17g %.17g %zu %zu |", p.fval, p.deltaFval, p.vars.size(), p.factors.size());
      for (size_t i = 0; i < p.vars.size(); ++i) std::printf(" %lld:%.17g", p.vars[i]->getID(), p.xval[i]);
      std::printf("\n");
    }
    // post-conditions: host variables hold the final values; the full objective is consistent
    for (const ComponentProblem& p : probs)
      for (size_t i = 0; i < p.vars.size(); ++i)
        if (!p.vars[i]->isAssigned() || p.vars[i]->eval() != p.xval[i]) {
          std::printf("POSTCONDITION VIOLATED var %lld\n", p.vars[i]->getID());
          return 1;
        }
    std::printf("eval_all %.17g\n", L.fn->eval());
    {  // interval bounds through the plugin surface: everything assigned -> the point value
      const NumericInterval b = L.fn->computeBounds();
      std::printf("bounds_all %.17g %.17g\n", b.lower(), b.upper());
    }
    // gradient of everything w.r.t. the first problem's variables, through the plugin surface
    if (!probs.empty()) {
      NumericVec g;
      L.fn->computeGradient(L.fn->getFactors(), probs[0].vars, g);
      std::printf("grad0");
      for (double v : g) std::printf(" %.17g", v);
      std::printf("\n");
    }
    return 0;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "host_driver: %s\n", e.what());
    return 1;
  } catch (const char* s) {
    std::fprintf(stderr, "host_driver: %s\n", s);
    return 1;
  }
}
