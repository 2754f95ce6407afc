// oracle/_ref/libguiding_ref.so — the reference's OWN guiding-fit code, compiled where it lies.
//
// TEST INFRASTRUCTURE ONLY.  Nothing in the product (rtx-pathtracer_b200/) includes, links or loads this file; it is
// used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg as the checker / CPU baseline.
//
// What is the reference's and what is restated here:
//   * external/lightpmm/include/pmm/*  (VMMFactory, VMFKernel, ParametricMixtureModel, fastexp, ...) and
//     external/guiding/*               (IncrementalDistance, IncrementalPearsonChiSquared, IncrementalCovariance2D)
//     are #included from /root/reference UNCHANGED (header-only, SSE 4-wide).  GLM, a system dependency that is not
//     installed here, is replaced by the ~60-line shim in oracle/glm_shim (vec3::length() == 3).
//   * src/PathGuiding.cpp cannot be compiled (Vulkan + Eigen), so the arithmetic of its update path is restated below,
//     function by function, with the reference line ranges in the comments:
//       createRegions / Aabb::addEpsilon / splitAabb   PathGuiding.cpp:81-104, Shapes.h:32-53
//       getSortedData                                  SampleCollector.cpp:76-131   (stable sort here: the reference's
//                                                      std::sort(par_unseq) leaves the order within a region unspecified)
//       update / updateRegion / preFit / postFit       PathGuiding.cpp:276-312, 350-451
//       computeEigenValuesVectors                      PathGuiding.cpp:454-484     (Eigen::EigenSolver restated for 2x2,
//                                                      see eigen2x2 below — Eigen is absent: this piece is pinned to the
//                                                      DEFINITION only, M v = lambda v against numpy in float64, tests/test_guiding_cpu.py)
//       splitComponentUsingPCA / splitAll              PathGuiding.cpp:494-633
//       mergeAll / computePearsonChiSquaredMergeMetric / mergeComponents   PathGuiding.cpp:635-788
//       pmmToVMM_Theta / syncPMMsToVMM_Thetas / VMF_Theta::setK            PathGuiding.cpp:53-69,106-132, PathGuiding.h:36-56
#include <pmm/VMMFactory.h>       // first, like src/PathGuiding.h:11 (it fixes the include order of the other pmm headers)
#include <pmm/DirectionalData.h>
#include <guiding/Range.h>
#include <guiding/incrementaldistance.h>
#include <guiding/incrementalpearsonchisquared.h>
#include <guiding/incrementalcovariance2d.h>

#include <algorithm>
#include <array>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <thread>
#include <tuple>
#include <vector>

#include "../include/b200pt.h"

namespace {

typedef lightpmm::Scalar4 Scalar;                                                    // PathGuiding.h:28-30
typedef lightpmm::VMFKernel<Scalar> VMF;
typedef lightpmm::ParametricMixtureModel<VMF, B200PT_MAX_DISTRIBUTIONS / Scalar::Width::value> PMM;
using lightpmm::DirectionalData;
static_assert(sizeof(DirectionalData) == sizeof(b200pt_directional_data), "DirectionalData layout");
typedef guiding::Range<std::vector<DirectionalData>> SampleRange;

struct ExtraData {                                                                   // PMM_ExtraData, PathGuiding.h:77-85
    glm::vec3 parallaxMean, lastParallaxMean;
    guiding::IncrementalDistance<PMM> incrementalDistance;
    guiding::IncrementalCovariance2D<PMM> incrementalCovariance2D;
    guiding::IncrementalPearsonChiSquared<PMM> incrementalPearsonChiSquared;
    uint32_t samplesSinceLastMerge = 0;
};

// ---- Eigen::EigenSolver<MatrixXf> on a 2x2 matrix, restated (Eigen 3.3 RealSchur::compute / splitOffTwoRows,
// JacobiRotation::makeGivens, EigenSolver::doComputeEigenvectors).  Eigen is a system dependency of the reference and
// is not available here, so this follows the published algorithm; only the 2x2 real-eigenvalue path is needed because
// the input is a symmetric covariance matrix.  V holds the normalised eigenvectors as COLUMNS, like Eigen.
void eigen2x2(const float m[2][2] /*row-major*/, float eval[2], float V[2][2] /*V[row][col]*/) {
    float T[2][2] = {{m[0][0], m[0][1]}, {m[1][0], m[1][1]}};
    float U[2][2] = {{1.0f, 0.0f}, {0.0f, 1.0f}};
    const float eps = std::numeric_limits<float>::epsilon();
    const float considerAsZero = std::numeric_limits<float>::min();
    float scale = std::max(std::max(std::fabs(T[0][0]), std::fabs(T[0][1])), std::max(std::fabs(T[1][0]), std::fabs(T[1][1])));
    if (scale < considerAsZero) {
        T[0][0] = T[0][1] = T[1][0] = T[1][1] = 0.0f;
    } else {
        for (auto &row : T) for (float &v : row) v /= scale;
        // findSmallSubdiagEntry
        float s = std::fabs(T[0][0]) + std::fabs(T[1][1]);
        s = std::max(s * eps, considerAsZero);
        if (std::fabs(T[1][0]) <= s) {
            T[1][0] = 0.0f;                                  // one root found; nothing to rotate
        } else {                                             // splitOffTwoRows(iu = 1)
            const float p = 0.5f * (T[0][0] - T[1][1]);
            const float q = p * p + T[1][0] * T[0][1];
            if (q >= 0.0f) {
                const float z = std::sqrt(std::fabs(q));
                const float gp = (p >= 0.0f) ? p + z : p - z, gq = T[1][0];
                float c, sn;                                 // makeGivens(gp, gq)
                if (gq == 0.0f) { c = gp < 0.0f ? -1.0f : 1.0f; sn = 0.0f; }
                else if (gp == 0.0f) { c = 0.0f; sn = gq < 0.0f ? 1.0f : -1.0f; }
                else if (std::fabs(gp) > std::fabs(gq)) {
                    const float t = gq / gp; float u = std::sqrt(1.0f + t * t); if (gp < 0.0f) u = -u;
                    c = 1.0f / u; sn = -t * c;
                } else {
                    const float t = gp / gq; float u = std::sqrt(1.0f + t * t); if (gq < 0.0f) u = -u;
                    sn = -1.0f / u; c = -t * sn;
                }
                // T.applyOnTheLeft(0, 1, rot.adjoint()): rows x=row0, y=row1 with (c, -s): x' = c x - s y ; y' = s x + c y
                for (int j = 0; j < 2; j++) { const float x = T[0][j], y = T[1][j]; T[0][j] = c * x - sn * y; T[1][j] = sn * x + c * y; }
                // T.applyOnTheRight(0, 1, rot): columns x=col0, y=col1 with rot.transpose() = (c, -s)
                for (int i = 0; i < 2; i++) { const float x = T[i][0], y = T[i][1]; T[i][0] = c * x - sn * y; T[i][1] = sn * x + c * y; }
                T[1][0] = 0.0f;
                for (int i = 0; i < 2; i++) { const float x = U[i][0], y = U[i][1]; U[i][0] = c * x - sn * y; U[i][1] = sn * x + c * y; }
            }
        }
        for (auto &row : T) for (float &v : row) v *= scale;
    }
    eval[0] = T[0][0]; eval[1] = T[1][1];
    // doComputeEigenvectors: back substitution on the (upper triangular) T, then multiply by U
    float norm = 0.0f;
    for (int j = 0; j < 2; j++) for (int i = std::max(j - 1, 0); i < 2; i++) norm += std::fabs(T[j][i]);
    float x01 = 0.0f;                                        // T(0,1) after back substitution for n = 1
    if (norm != 0.0f) {
        const float w = T[0][0] - eval[1];
        const float r = T[0][1];                             // T.row(0).segment(1,1) . T.col(1).segment(1,1) with T(1,1) = 1
        x01 = (w != 0.0f) ? -r / w : -r / (eps * norm);
        // (Eigen's overflow control would rescale this column by 1/abs(x01); the column is normalised below anyway)
    }
    float col1[2] = {U[0][0] * x01 + U[0][1] * 1.0f, U[1][0] * x01 + U[1][1] * 1.0f};
    float col0[2] = {U[0][0], U[1][0]};
    const float n0 = std::sqrt(col0[0] * col0[0] + col0[1] * col0[1]), n1 = std::sqrt(col1[0] * col1[0] + col1[1] * col1[1]);
    V[0][0] = col0[0] / n0; V[1][0] = col0[1] / n0;
    V[0][1] = col1[0] / n1; V[1][1] = col1[1] / n1;
}

// PathGuiding.cpp:454-484 — note the ROWS of the eigenvector matrix are paired with the eigenvalues (quirk 9), and
// the sort comparator is `<=` (on two elements libstdc++'s insertion sort swaps them when eval[1] <= eval[0])
void computeEigenValuesVectors(const lightpmm::Matrix2x2 &covmat, float eigenValues[2], glm::vec2 eigenVectors[2]) {
    const float mat[2][2] = {{covmat[0][0], covmat[1][0]}, {covmat[0][1], covmat[1][1]}};
    float ev[2], V[2][2];
    eigen2x2(mat, ev, V);
    int order[2] = {0, 1};
    if (ev[1] <= ev[0]) { order[0] = 1; order[1] = 0; }
    for (int k = 0; k < 2; k++) {
        eigenValues[k] = ev[order[k]];
        eigenVectors[k] = glm::vec2(V[order[k]][0], V[order[k]][1]);     // eigen_vectors.row(i)
    }
}

static void splitAabb(const b200pt_aabb &b, b200pt_aabb &l, b200pt_aabb &r) {        // Aabb::splitAabb, Shapes.h:32-46
    const float size[3] = {b.max[0] - b.min[0], b.max[1] - b.min[1], b.max[2] - b.min[2]};
    const int axis = size[0] > size[1] ? (size[0] > size[2] ? 0 : 2) : (size[1] > size[2] ? 1 : 2);
    l = b; r = b;
    l.max[axis] -= 0.5f * size[axis];
    r.min[axis] += 0.5f * size[axis];
}

struct RefGuiding {
    uint32_t regionCount = 0;
    bool firstFit = true;
    bool useParallaxCompensation = true;
    b200pt_guiding_params gp{};
    lightpmm::VMMFactoryProperties props{};
    lightpmm::VMMFactory<PMM> vmmFactory;
    std::vector<b200pt_aabb> aabbs;
    std::vector<PMM> pmms;
    std::vector<ExtraData> extra;
    std::vector<b200pt_vmm_theta> thetas;
    std::vector<DirectionalData> sorted;
    std::vector<uint32_t> offsets;
    uint64_t emSampleIterations = 0;    // sum over regions of N_r * EM iterations of fit/updateFit (bench accounting)

    void configure(const b200pt_guiding_params &p) {                                 // PathGuiding.cpp:33-40
        gp = p;
        props.numInitialComponents = uint32_t(p.numInitialComponents);
        props.minItr = uint32_t(p.minItr); props.maxItr = uint32_t(p.maxItr);
        props.relLogLikelihoodThreshold = p.relLogLikelihoodThreshold;
        props.initKappa = p.initKappa; props.maxKappa = p.maxKappa;
        props.vPrior = p.vPrior; props.rPrior = p.rPrior; props.rPriorWeight = p.rPriorWeight;
        vmmFactory = lightpmm::VMMFactory<PMM>(props);
        useParallaxCompensation = p.useParallaxCompensation != 0;
    }

    void createRegions(int splits, const float smin[3], const float smax[3]) {       // PathGuiding.cpp:81-104
        b200pt_aabb scene;
        for (int a = 0; a < 3; a++) {                                                // Aabb::addEpsilon, Shapes.h:48-53
            const float extent = smax[a] - smin[a];
            const float center = smin[a] + 0.5f * extent;
            scene.min[a] = center - 0.50001f * extent;
            scene.max[a] = center + 0.50001f * extent;
        }
        aabbs.assign(1, scene);
        for (int i = 0; i < splits; i++) {
            std::vector<b200pt_aabb> next;
            for (const b200pt_aabb &b : aabbs) { b200pt_aabb l, r; splitAabb(b, l, r); next.push_back(l); next.push_back(r); }
            aabbs.swap(next);
        }
        regionCount = uint32_t(aabbs.size());
        pmms.assign(regionCount, PMM());                                             // createPMMs, PathGuiding.cpp:42-51
        extra.assign(regionCount, ExtraData());
        for (PMM &pmm : pmms) vmmFactory.initialize(pmm);
        firstFit = true;
        syncThetas();
    }

    static void setK(b200pt_vmf_theta &t, float newK) {                              // VMF_Theta::setK, PathGuiding.h:45-50 (double math)
        newK = newK < VMF_MinKappa ? 0.0 : newK;
        t.k = newK;
        t.norm = t.k / (2 * M_PI * (1 - exp(-2 * t.k)));
        t.eMin2K = exp(-2.0 * t.k);
    }

    void syncThetas() {                                                              // PathGuiding.cpp:53-69, 106-132
        thetas.assign(regionCount, b200pt_vmm_theta());
        for (uint32_t r = 0; r < regionCount; r++) {
            b200pt_vmm_theta &vmm = thetas[r];
            memset(&vmm, 0, sizeof(vmm));
            for (auto &t : vmm.thetas) t.distance = -1.0f;
            const PMM &pmm = pmms[r];
            vmm.usedDistributions = int(pmm.getK());
            for (int i = 0; i < vmm.usedDistributions; i++) {
                vmm.pi[i] = pmm.weightK(i);
                const glm::vec3 mu = pmm.m_comps[i / 4].getMu(i % 4);
                vmm.thetas[i].mu[0] = mu.x; vmm.thetas[i].mu[1] = mu.y; vmm.thetas[i].mu[2] = mu.z;
                setK(vmm.thetas[i], pmm.m_comps[i / 4].getKappa(i % 4));
            }
            if (useParallaxCompensation) {
                const glm::vec3 mean = extra[r].parallaxMean;
                vmm.meanPosition[0] = mean.x; vmm.meanPosition[1] = mean.y; vmm.meanPosition[2] = mean.z;
                for (int i = 0; i < B200PT_MAX_DISTRIBUTIONS; i++) {
                    float distance = extra[r].incrementalDistance.distances[i / 4][i % 4];
                    distance = distance == std::numeric_limits<float>::infinity() ? -1.0f : distance;
                    vmm.thetas[i].distance = distance;
                    for (int a = 0; a < 3; a++) vmm.thetas[i].target[a] = vmm.meanPosition[a] + distance * vmm.thetas[i].mu[a];
                }
            }
        }
    }

    // SampleCollector::getSortedData (SampleCollector.cpp:76-131): order by region id, INVALID last; region offsets
    void sortSamples(const b200pt_directional_data *in, int64_t n) {
        std::vector<uint32_t> count(regionCount + 1, 0);
        for (int64_t i = 0; i < n; i++) if (in[i].flags < regionCount) count[in[i].flags + 1]++;
        offsets.assign(regionCount + 1, 0);
        for (uint32_t r = 0; r < regionCount; r++) offsets[r + 1] = offsets[r] + count[r + 1];
        sorted.resize(offsets[regionCount]);
        std::vector<uint32_t> cursor(offsets.begin(), offsets.end() - 1);
        for (int64_t i = 0; i < n; i++)
            if (in[i].flags < regionCount) memcpy(&sorted[cursor[in[i].flags]++], &in[i], sizeof(DirectionalData));
    }

    void preFit(uint32_t iRegion, uint32_t begin, uint32_t end, bool isFirst) {       // PathGuiding.cpp:370-411
        if (!useParallaxCompensation) return;
        const b200pt_aabb &bb = aabbs[iRegion];
        const glm::vec3 mn(bb.min[0], bb.min[1], bb.min[2]), mx(bb.max[0], bb.max[1], bb.max[2]);
        const glm::vec3 parallaxMean = mn + 0.5f * (mx - mn);
        extra[iRegion].lastParallaxMean = extra[iRegion].parallaxMean;
        extra[iRegion].parallaxMean = parallaxMean;
        for (uint32_t i = begin; i < end; i++) {
            DirectionalData &s = sorted[i];
            if (s.distance > 0.0f) {
                const glm::vec3 nd = s.position + s.distance * s.direction - parallaxMean;
                s.direction = glm::normalize(nd);
                s.distance = glm::length(nd);
            } else {
                s.distance = std::numeric_limits<float>::infinity();
            }
            s.position = parallaxMean;
        }
        if (!isFirst) extra[iRegion].incrementalDistance.reposition(pmms[iRegion], extra[iRegion].lastParallaxMean - parallaxMean);
    }

    void splitComponentUsingPCA(ExtraData &stats, PMM &pmm, uint32_t component, float maxKappa) {   // PathGuiding.cpp:494-561
        if (pmm.getK() == PMM::MaxK::value) return;
        const uint32_t sourceIndex = component, targetIndex = pmm.getK();
        VMF &sourceComponent = pmm.getComponent(sourceIndex / 4);
        VMF &targetComponent = pmm.getComponent(targetIndex / 4);
        const lightpmm::Frame frame{sourceComponent.getMu(component % 4)};
        const lightpmm::Matrix2x2 covariance = stats.incrementalCovariance2D.computeCovarianceMatrix(component);
        float eigenValues[2]; glm::vec2 eigenVectors[2];
        computeEigenValuesVectors(covariance, eigenValues, eigenVectors);
        const int maxEigValIndex = eigenValues[1] > eigenValues[0];
        const float oneHalfSigmaOffset = std::min(1.0f, 0.5f * std::sqrt(eigenValues[maxEigValIndex]));
        lightpmm::Vector2 principalComponentDir = eigenVectors[maxEigValIndex];
        principalComponentDir *= oneHalfSigmaOffset;
        const float z = std::sqrt(1.0f - oneHalfSigmaOffset * oneHalfSigmaOffset);
        const lightpmm::Vector3 splitMuA{frame.toWorld({principalComponentDir.x, principalComponentDir.y, z})};
        const lightpmm::Vector3 splitMuB{frame.toWorld({-principalComponentDir.x, -principalComponentDir.y, z})};
        pmm.setK(targetIndex + 1);
        const float sourceAvgCosine = sourceComponent.getR()[sourceIndex % 4];
        const float sourceWeight = sourceComponent.getWeight(sourceIndex % 4);
        const float splitWeight = sourceWeight * 0.5f;
        const float maxAvgCosine = lightpmm::kappaToMeanCosine(maxKappa);
        const float splitAvgCosine = (z > 0.0f) ? std::min(maxAvgCosine, sourceAvgCosine / z) : maxAvgCosine;
        const float splitKappa = lightpmm::meanCosineToKappa(splitAvgCosine);
        sourceComponent.setKappaAndR(sourceIndex % 4, splitKappa, splitAvgCosine);
        sourceComponent.setMu(sourceIndex % 4, splitMuA);
        sourceComponent.setWeight(sourceIndex % 4, splitWeight);
        targetComponent.setKappaAndR(targetIndex % 4, splitKappa, splitAvgCosine);
        targetComponent.setMu(targetIndex % 4, splitMuB);
        targetComponent.setWeight(targetIndex % 4, splitWeight);
        stats.incrementalDistance.split(pmm, component);
        stats.incrementalPearsonChiSquared.split(pmm, component);
        stats.incrementalCovariance2D.split(pmm, component);
    }

    uint32_t splitAll(ExtraData &stats, PMM &pmm, const SampleRange &samples, bool iterative, bool fit) {   // PathGuiding.cpp:563-633
        if (pmm.getK() >= PMM::MaxK::value) return 0;
        const bool firstFitLocal = samples.size() == pmm.m_totalNumSamples;
        uint32_t totalNumSplits = 0, numSplits = 0;
        do {
            const uint32_t numActiveKernels = (pmm.getK() + 3) / 4;
            const std::array<Scalar, PMM::NumKernels::value> divergence = stats.incrementalPearsonChiSquared.computeDivergence(pmm);
            std::array<std::pair<float, uint32_t>, PMM::MaxK::value> splitCandidates;
            for (uint32_t k = 0; k < numActiveKernels; ++k) {
                const Scalar::BooleanType seenEnoughSamples = stats.incrementalPearsonChiSquared.numSamples[k] > float(gp.minSamplesForSplitting);
                const Scalar weightedDivergence = lightpmm::ifthen(firstFitLocal || seenEnoughSamples,
                                                                   divergence[k] * pmm.getComponent(k).m_weights, 0.0f);
                for (uint32_t i = 0; i < 4; ++i) splitCandidates[k * 4 + i] = std::make_pair(std::fabs(weightedDivergence[i]), k * 4 + i);
            }
            const float splitMinDivergence = gp.splitMinDivergence;
            const auto lastCandidateIterator = std::partition(splitCandidates.begin(), splitCandidates.begin() + pmm.getK(),
                [splitMinDivergence](const std::pair<float, uint32_t> d) -> bool { return d.first >= splitMinDivergence; });
            const uint32_t numSplitCandidates = uint32_t(std::distance(splitCandidates.begin(), lastCandidateIterator));
            const uint32_t maxNumSplits = PMM::MaxK::value - pmm.getK();
            numSplits = std::min(numSplitCandidates, maxNumSplits);
            if (numSplits == 0) break;
            if (numSplitCandidates > numSplits)
                std::partial_sort(splitCandidates.begin(), splitCandidates.begin() + numSplits, lastCandidateIterator,
                                  [](const std::pair<float, uint32_t> a, const std::pair<float, uint32_t> b) -> bool { return a.first > b.first; });
            std::array<Scalar::BooleanType, PMM::NumKernels::value> modifiedComponentMask;
            std::fill(modifiedComponentMask.begin(), modifiedComponentMask.end(), Scalar::BooleanType{false});
            for (uint32_t i = 0; i < numSplits; ++i) {
                modifiedComponentMask[splitCandidates[i].second / 4].insert(splitCandidates[i].second % 4, true);
                modifiedComponentMask[pmm.getK() / 4].insert(pmm.getK() % 4, true);
                splitComponentUsingPCA(stats, pmm, splitCandidates[i].second, vmmFactory.m_properties.maxKappa);
            }
            if (fit) {
                pmm.applyWeightPrior(props.vPrior);
                vmmFactory.maskedFit(samples.begin(), samples.end(), pmm, modifiedComponentMask);
                stats.incrementalPearsonChiSquared.updateDivergenceMasked(pmm, samples, modifiedComponentMask);
                stats.incrementalCovariance2D.updateStatisticsMasked(pmm, samples, modifiedComponentMask);
                pmm.removeWeightPrior(props.vPrior);
            }
            totalNumSplits += numSplits;
        } while (iterative);
        return totalNumSplits;
    }

    static std::array<float, 120> computePearsonChiSquaredMergeMetric(const PMM &distribution) {    // PathGuiding.cpp:713-767
        const uint32_t numComponents = distribution.getK();
        const uint32_t numActiveKernels = (numComponents + 3) / 4;
        std::array<VMF, PMM::NumKernels::value> selfProduct;
        for (uint32_t k = 0; k < numActiveKernels; ++k) { selfProduct[k] = distribution.getComponent(k); selfProduct[k].product(selfProduct[k]); }
        std::array<Scalar, PMM::NumKernels::value * PMM::MaxK::value> sim;
        for (uint32_t i = 0; i < numComponents - 1; ++i) {
            const VMF componentI = distribution.getComponent(i / 4).extract(i % 4);
            const VMF componentISqr = selfProduct[i / 4].extract(i % 4);
            for (uint32_t j = (i + 1) / 4; j < numActiveKernels; ++j) {
                VMF productIJ{componentI};
                productIJ.product(distribution.getComponent(j));
                const VMF kernelJ{distribution.getComponent(j)};
                VMF merged{componentI};
                { VMF kJForMerge{kernelJ}; for (uint32_t l = 0; l < 4; ++l) merged.mergeComponent(l, l, kJForMerge); }
                const Scalar quotientISqrMerged = VMF{componentISqr}.division(merged);
                const Scalar quotientIJMerged = VMF{productIJ}.division(merged);
                const Scalar quotientJSqrMerged = VMF{selfProduct[j]}.division(merged);
                sim[i * PMM::NumKernels::value + j] = quotientISqrMerged + 2.0f * quotientIJMerged + quotientJSqrMerged - merged.m_weights;
            }
        }
        std::array<float, 120> out{};
        uint32_t index = 0;
        for (uint32_t i = 0; i < numComponents - 1; ++i)
            for (size_t j = i + 1; j < numComponents; ++j, ++index) out[index] = sim[i * PMM::NumKernels::value + j / 4][j % 4];
        return out;
    }

    void mergeComponents(uint32_t iRegion, uint32_t a, uint32_t b) {                   // PathGuiding.cpp:769-788
        PMM &pmm = pmms[iRegion];
        extra[iRegion].incrementalDistance.merge(pmm, a, b);
        extra[iRegion].incrementalPearsonChiSquared.merge(pmm, a, b);
        extra[iRegion].incrementalCovariance2D.merge(pmm, a, b);
        pmm.mergeComponents(a, b);
    }

    uint32_t mergeAll(uint32_t iRegion) {                                              // PathGuiding.cpp:635-711
        uint32_t totalNumMerges = 0;
        PMM &pmm = pmms[iRegion];
        do {
            const uint32_t numComponents = pmm.getK();
            uint32_t numMerges = 0;
            if (numComponents <= 1) return 0;
            const uint32_t numSimilarityValues = numComponents * (numComponents - 1) / 2;
            const std::array<float, 120> mergeMetric = computePearsonChiSquaredMergeMetric(pmm);
            std::array<std::pair<float, std::pair<uint32_t, uint32_t>>, 120> cand;
            for (uint32_t a = 0, offA = 0; a < numComponents; ++a, offA += numComponents - a)
                for (uint32_t b = a + 1, offB = 0; b < numComponents; ++b, ++offB)
                    cand[offA + offB] = std::make_pair(mergeMetric[offA + offB], std::make_pair(a, b));
            const float mergeMaxDivergence = gp.mergeMaxDivergence;
            const auto validEnd = std::partition(cand.begin(), cand.begin() + numSimilarityValues,
                [mergeMaxDivergence](std::pair<float, std::pair<uint32_t, uint32_t>> c) -> bool { return c.first <= mergeMaxDivergence; });
            std::sort(cand.begin(), validEnd, [](std::pair<float, std::pair<uint32_t, uint32_t>> x, std::pair<float, std::pair<uint32_t, uint32_t>> y) -> bool { return x.first < y.first; });
            uint32_t bitmask = 0;
            std::array<std::pair<uint32_t, uint32_t>, PMM::MaxK::value> merges;
            for (auto it = cand.begin(); it != validEnd; ++it) {
                const uint32_t a = std::min(it->second.first, it->second.second), b = std::max(it->second.first, it->second.second);
                if ((bitmask & (1u << a)) || (bitmask & (1u << b))) continue;
                merges[numMerges++] = std::make_pair(a, b);
                bitmask |= (1u << a) | (1u << b);
            }
            if (numMerges == 0) break;
            std::sort(merges.begin(), merges.begin() + numMerges, [](std::pair<uint32_t, uint32_t> x, std::pair<uint32_t, uint32_t> y) -> bool { return x.second > y.second; });
            for (uint32_t i = 0; i < numMerges; ++i) mergeComponents(iRegion, merges[i].first, merges[i].second);
            totalNumMerges += numMerges;
        } while (true);
        return totalNumMerges;
    }

    void postFit(uint32_t iRegion, const SampleRange &range) {                         // PathGuiding.cpp:413-451
        PMM &pmm = pmms[iRegion];
        ExtraData &x = extra[iRegion];
        if (gp.splitAndMerge) {
            x.samplesSinceLastMerge += uint32_t(range.size());
            if (x.samplesSinceLastMerge > uint32_t(gp.minSamplesForMerging)) {
                pmm.removeWeightPrior(props.vPrior);
                mergeAll(iRegion);
                pmm.applyWeightPrior(props.vPrior);
                x.samplesSinceLastMerge = 0;
            }
            x.incrementalPearsonChiSquared.updateDivergence(pmm, range);
            x.incrementalCovariance2D.updateStatistics(pmm, range);
            const bool firstFitLocal = range.size() == pmm.m_totalNumSamples;
            const bool fitAfterSplit = firstFitLocal || range.size() > size_t(gp.minSamplesForPostSplitFitting);
            pmm.removeWeightPrior(props.vPrior);
            splitAll(x, pmm, range, true, fitAfterSplit);
            pmm.applyWeightPrior(props.vPrior);
        }
        if (useParallaxCompensation) x.incrementalDistance.updateDistances(pmm, range);
    }

    void updateRegion(uint32_t iRegion, uint32_t begin, uint32_t end, bool isFirst) {  // PathGuiding.cpp:350-368
        SampleRange range(sorted.begin() + begin, sorted.begin() + end);
        preFit(iRegion, begin, end, isFirst);
        const uint32_t before = pmms[iRegion].m_numEMIterations;
        if (isFirst) vmmFactory.fit(range.begin(), range.end(), pmms[iRegion], false);
        else vmmFactory.updateFit(range.begin(), range.end(), pmms[iRegion]);
        __atomic_fetch_add(&emSampleIterations, uint64_t(end - begin) * (pmms[iRegion].m_numEMIterations - before), __ATOMIC_RELAXED);
        postFit(iRegion, range);
    }

    // PathGuiding::update (PathGuiding.cpp:276-312).  threads == 1 is the reference's serial region loop; threads > 1
    // runs the same per-region code on a thread pool (regions are independent), the "all host cores" baseline.
    void update(const b200pt_directional_data *in, int64_t n, int threads) {
        sortSamples(in, n);
        const bool isFirst = firstFit;
        if (threads <= 1) {
            for (uint32_t r = 0; r < regionCount; r++) if (offsets[r + 1] - offsets[r] > 0) updateRegion(r, offsets[r], offsets[r + 1], isFirst);
        } else {
            std::atomic<uint32_t> next{0};
            std::vector<std::thread> pool;
            for (int t = 0; t < threads; t++)
                pool.emplace_back([&]() {
                    for (;;) {
                        const uint32_t r = next.fetch_add(1);
                        if (r >= regionCount) break;
                        if (offsets[r + 1] - offsets[r] > 0) updateRegion(r, offsets[r], offsets[r + 1], isFirst);
                    }
                });
            for (auto &t : pool) t.join();
        }
        firstFit = false;
        // Check for regions to split (PathGuiding.cpp:291-300) + splitRegion (:328-348)
        const uint32_t currentRegionCount = regionCount;
        for (uint32_t r = 0; r < currentRegionCount; r++) {
            if (gp.splitRegions && pmms[r].m_numSamples > gp.samplesForRegionSplit) {
                b200pt_aabb left, right;
                splitAabb(aabbs[r], left, right);
                aabbs[r] = left;
                aabbs.push_back(right);
                const float decayTerm = 0.25f;
                pmms[r].m_numSamples *= decayTerm;
                pmms[r].m_sampleWeight *= decayTerm;
                pmms.push_back(pmms[r]);
                extra.push_back(extra[r]);
                regionCount++;
            }
        }
        syncThetas();
    }
};

}  // namespace

extern "C" {

void *refguiding_create(int splits, const float scene_min[3], const float scene_max[3], const b200pt_guiding_params *params) {
    RefGuiding *g = new RefGuiding();
    g->configure(*params);
    g->createRegions(splits, scene_min, scene_max);
    return g;
}
void refguiding_destroy(void *h) { delete static_cast<RefGuiding *>(h); }
// the restated 2x2 eigen decomposition on its own (row-major m, eigenvectors as columns of V), for tests/test_guiding_cpu.py
void refguiding_eigen2x2(const float m[4], float eval[2], float V[4]) {
    const float mm[2][2] = {{m[0], m[1]}, {m[2], m[3]}};
    float vv[2][2];
    eigen2x2(mm, eval, vv);
    V[0] = vv[0][0]; V[1] = vv[0][1]; V[2] = vv[1][0]; V[3] = vv[1][1];
}
int refguiding_region_count(void *h) { return int(static_cast<RefGuiding *>(h)->regionCount); }
void refguiding_get_aabbs(void *h, b200pt_aabb *out) { auto *g = static_cast<RefGuiding *>(h); memcpy(out, g->aabbs.data(), g->aabbs.size() * sizeof(b200pt_aabb)); }
void refguiding_update(void *h, const b200pt_directional_data *samples, int64_t n, int threads) { static_cast<RefGuiding *>(h)->update(samples, n, threads); }
void refguiding_get_vmms(void *h, b200pt_vmm_theta *out) { auto *g = static_cast<RefGuiding *>(h); memcpy(out, g->thetas.data(), g->thetas.size() * sizeof(b200pt_vmm_theta)); }
uint64_t refguiding_em_sample_iterations(void *h) { return static_cast<RefGuiding *>(h)->emSampleIterations; }
int64_t refguiding_sorted_count(void *h) { return int64_t(static_cast<RefGuiding *>(h)->sorted.size()); }
// sorted (and pre-fitted) samples of the last update + region offsets [regionCount + 1]
void refguiding_get_sorted(void *h, b200pt_directional_data *out, uint32_t *offsets) {
    auto *g = static_cast<RefGuiding *>(h);
    if (out) memcpy(out, g->sorted.data(), g->sorted.size() * sizeof(DirectionalData));
    if (offsets) memcpy(offsets, g->offsets.data(), g->offsets.size() * sizeof(uint32_t));
}
// full per-region state for deep parity checks: K, sampleWeight, numSamples, totalNumSamples, numEMIterations and the
// per-component arrays (16 each): weight, kappa, r, mu xyz, distance, distance sumWeights, chi2 value, chi2 numSamples,
// cov xx yy xy, cov sumWeights
void refguiding_get_state(void *h, int region, float *scalars5, float *perComponent /* 14 x 16 */) {
    auto *g = static_cast<RefGuiding *>(h);
    const PMM &pmm = g->pmms[region];
    const ExtraData &x = g->extra[region];
    scalars5[0] = float(pmm.getK()); scalars5[1] = pmm.m_sampleWeight; scalars5[2] = pmm.m_numSamples;
    scalars5[3] = float(pmm.m_totalNumSamples); scalars5[4] = float(pmm.m_numEMIterations);
    for (int i = 0; i < 16; i++) {
        const VMF &c = pmm.m_comps[i / 4];
        const int l = i % 4;
        float *p = perComponent;
        p[0 * 16 + i] = c.m_weights[l]; p[1 * 16 + i] = c.getKappa()[l]; p[2 * 16 + i] = c.getR()[l];
        p[3 * 16 + i] = c.m_mu.x[l]; p[4 * 16 + i] = c.m_mu.y[l]; p[5 * 16 + i] = c.m_mu.z[l];
        p[6 * 16 + i] = x.incrementalDistance.distances[i / 4][l]; p[7 * 16 + i] = x.incrementalDistance.sumWeights[i / 4][l];
        p[8 * 16 + i] = x.incrementalPearsonChiSquared.divergencePlusOneTimesIntegralSqr[i / 4][l];
        p[9 * 16 + i] = x.incrementalPearsonChiSquared.numSamples[i / 4][l];
        p[10 * 16 + i] = x.incrementalCovariance2D.varianceAndCovariance[i / 4].x[l];
        p[11 * 16 + i] = x.incrementalCovariance2D.varianceAndCovariance[i / 4].y[l];
        p[12 * 16 + i] = x.incrementalCovariance2D.varianceAndCovariance[i / 4].z[l];
        p[13 * 16 + i] = x.incrementalCovariance2D.sumWeights[i / 4][l];
    }
}
// lightpmm::exp (PMM_APPROX_EXP fastexp) on a float array: known-answer hook for the device fastexp
void refguiding_fastexp(const float *in, float *out, int n) {
    for (int i = 0; i < n; i += 4) {
        float tmp[4] = {0, 0, 0, 0};
        for (int k = 0; k < 4 && i + k < n; k++) tmp[k] = in[i + k];
        lightpmm::float4 v; v.load(tmp);
        lightpmm::float4 e = lightpmm::exp(v);
        e.store(tmp);
        for (int k = 0; k < 4 && i + k < n; k++) out[i + k] = tmp[k];
    }
}
}  // extern "C"
