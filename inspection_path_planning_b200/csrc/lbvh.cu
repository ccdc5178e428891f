// lbvh.cu -- K1: GPU BVH build over the inspection mesh.
// 63-bit Morton codes of the centroids -> radix sort (CUB) -> Karras binary radix tree -> bottom-up
// refit with arrival counters -> 64-byte two-child nodes, subtrees of <= kLeafMax triangles collapsed
// into leaves, triangles gathered into BVH order as float4 SoA arrays.
// The reference has no acceleration structure on this path; this is new work.
#include <cub/cub.cuh>

#include <vector>

#include "kernels.h"

namespace ipp {

namespace {

__device__ __forceinline__ unsigned long long expand21(unsigned long long v) {
    v &= 0x1fffffull;
    v = (v | v << 32) & 0x1f00000000ffffull;
    v = (v | v << 16) & 0x1f0000ff0000ffull;
    v = (v | v << 8) & 0x100f00f00f00f00full;
    v = (v | v << 4) & 0x10c30c30c30c30c3ull;
    v = (v | v << 2) & 0x1249249249249249ull;
    return v;
}

__global__ void k_morton(const float* __restrict__ centroid, int T, float3 lo, float3 invExt, unsigned long long* __restrict__ keys,
                         int* __restrict__ vals) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T) return;
    float x = (centroid[3 * i] - lo.x) * invExt.x, y = (centroid[3 * i + 1] - lo.y) * invExt.y, z = (centroid[3 * i + 2] - lo.z) * invExt.z;
    x = fminf(fmaxf(x, 0.f), 1.f);
    y = fminf(fmaxf(y, 0.f), 1.f);
    z = fminf(fmaxf(z, 0.f), 1.f);
    unsigned long long xi = (unsigned long long)(x * 2097151.f), yi = (unsigned long long)(y * 2097151.f),
                       zi = (unsigned long long)(z * 2097151.f);
    keys[i] = (expand21(xi) << 2) | (expand21(yi) << 1) | expand21(zi);
    vals[i] = i;
}

__device__ __forceinline__ int delta(const unsigned long long* __restrict__ k, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    unsigned long long a = k[i], b = k[j];
    if (a == b) return 64 + __clz(i ^ j);
    return __clzll(a ^ b);
}

// child links during the build: bit 31 set = leaf
__global__ void k_karras(const unsigned long long* __restrict__ keys, int n, int* __restrict__ left, int* __restrict__ right,
                         int* __restrict__ rlo, int* __restrict__ rhi, int* __restrict__ parentI, int* __restrict__ parentL) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2)
        if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = delta(keys, n, i, j);
    int s = 0;
    int t = l;
    do {
        t = (t + 1) >> 1;
        if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    int gamma = i + s * d + min(d, 0);
    int lo = min(i, j), hi = max(i, j);
    int L, R;
    if (lo == gamma) { L = gamma | 0x80000000; parentL[gamma] = i; } else { L = gamma; parentI[gamma] = i; }
    if (hi == gamma + 1) { R = (gamma + 1) | 0x80000000; parentL[gamma + 1] = i; } else { R = gamma + 1; parentI[gamma + 1] = i; }
    left[i] = L;
    right[i] = R;
    rlo[i] = lo;
    rhi[i] = hi;
    if (i == 0) parentI[0] = -1;
}

__global__ void k_refit(const float* __restrict__ verts, const int* __restrict__ vals, int n, const int* __restrict__ left,
                        const int* __restrict__ right, const int* __restrict__ parentI, const int* __restrict__ parentL,
                        float4* leafMin, float4* leafMax, float4* nodeMin, float4* nodeMax, int* arrive) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* p = verts + (size_t)vals[i] * 9;
    float4 mn = make_float4(fminf(p[0], fminf(p[3], p[6])), fminf(p[1], fminf(p[4], p[7])), fminf(p[2], fminf(p[5], p[8])), 0.f);
    float4 mx = make_float4(fmaxf(p[0], fmaxf(p[3], p[6])), fmaxf(p[1], fmaxf(p[4], p[7])), fmaxf(p[2], fmaxf(p[5], p[8])), 0.f);
    leafMin[i] = mn;
    leafMax[i] = mx;
    if (n == 1) return;
    int node = parentL[i];
    while (node >= 0) {
        __threadfence();
        if (atomicAdd(arrive + node, 1) == 0) return;  // the second arrival continues upwards
        int L = left[node], R = right[node];
        // L2 reads (__ldcg): the values were published by other threads before their atomicAdd
        float4 amn = __ldcg((L < 0) ? leafMin + (L & 0x7fffffff) : nodeMin + L), amx = __ldcg((L < 0) ? leafMax + (L & 0x7fffffff) : nodeMax + L);
        float4 bmn = __ldcg((R < 0) ? leafMin + (R & 0x7fffffff) : nodeMin + R), bmx = __ldcg((R < 0) ? leafMax + (R & 0x7fffffff) : nodeMax + R);
        nodeMin[node] = make_float4(fminf(amn.x, bmn.x), fminf(amn.y, bmn.y), fminf(amn.z, bmn.z), 0.f);
        nodeMax[node] = make_float4(fmaxf(amx.x, bmx.x), fmaxf(amx.y, bmx.y), fmaxf(amx.z, bmx.z), 0.f);
        node = parentI[node];
    }
}

__global__ void k_emit(int n, const int* __restrict__ left, const int* __restrict__ right, const int* __restrict__ rlo,
                       const int* __restrict__ rhi, const float4* __restrict__ leafMin, const float4* __restrict__ leafMax,
                       const float4* __restrict__ nodeMin, const float4* __restrict__ nodeMax, float pad, BvhNode* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    int link[2];
    float4 mn[2], mx[2];
    int ch[2] = {left[i], right[i]};
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        int L = ch[c];
        if (L < 0) {
            int idx = L & 0x7fffffff;
            link[c] = ~((idx << kLeafShift) | 1);
            mn[c] = leafMin[idx];
            mx[c] = leafMax[idx];
        } else {
            int cnt = rhi[L] - rlo[L] + 1;
            link[c] = (cnt <= kLeafMax) ? ~((rlo[L] << kLeafShift) | cnt) : L;
            mn[c] = nodeMin[L];
            mx[c] = nodeMax[L];
        }
        mn[c].x -= pad; mn[c].y -= pad; mn[c].z -= pad;
        mx[c].x += pad; mx[c].y += pad; mx[c].z += pad;
    }
    BvhNode nd;
    nd.q0 = make_float4(mn[0].x, mn[0].y, mn[0].z, mx[0].x);
    nd.q1 = make_float4(mx[0].y, mx[0].z, mn[1].x, mn[1].y);
    nd.q2 = make_float4(mn[1].z, mx[1].x, mx[1].y, mx[1].z);
    nd.q3 = make_float4(__int_as_float(link[0]), __int_as_float(link[1]), 0.f, 0.f);
    out[i] = nd;
}

// Cluster table of the binned visibility kernel: the maximal subtrees of at most C triangles (C = kLeafMax: exactly the leaves of
// the node array above; larger C for meshes with more leaves than a camera pose can sort), each a contiguous run of the
// BVH-ordered triangle arrays.  Clusters are the children of the nodes that hold more than C triangles (and of the root).
__device__ __forceinline__ bool cluster_child(int link, int C, const int* __restrict__ rlo, const int* __restrict__ rhi, int& first, int& cnt) {
    if (link < 0) {  // a single triangle
        first = link & 0x7fffffff;
        cnt = 1;
        return true;
    }
    first = rlo[link];
    cnt = rhi[link] - rlo[link] + 1;
    return cnt <= C;
}
__global__ void k_cluster_count(int n, int C, const int* __restrict__ left, const int* __restrict__ right, const int* __restrict__ rlo,
                                const int* __restrict__ rhi, int* __restrict__ counts) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const bool live = i == 0 || rhi[i] - rlo[i] + 1 > C;
    int c = 0, f, k;
    if (live) c = (int)cluster_child(left[i], C, rlo, rhi, f, k) + (int)cluster_child(right[i], C, rlo, rhi, f, k);
    counts[i] = c;
}
__global__ void k_cluster_fill(int n, int C, const int* __restrict__ left, const int* __restrict__ right, const int* __restrict__ rlo,
                               const int* __restrict__ rhi, const float4* __restrict__ leafMin, const float4* __restrict__ leafMax,
                               const float4* __restrict__ nodeMin, const float4* __restrict__ nodeMax, float pad,
                               const int* __restrict__ counts, const int* __restrict__ offsets, float4* __restrict__ clusterLo,
                               float4* __restrict__ clusterHi) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1 || counts[i] == 0) return;
    int o = offsets[i];
    const int ch[2] = {left[i], right[i]};
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        int first, cnt;
        if (!cluster_child(ch[c], C, rlo, rhi, first, cnt)) continue;
        const int L = ch[c];
        float4 mn = (L < 0) ? leafMin[L & 0x7fffffff] : nodeMin[L], mx = (L < 0) ? leafMax[L & 0x7fffffff] : nodeMax[L];
        // (the padding of k_emit: the same boxes as in the node array)
        mn.x -= pad; mn.y -= pad; mn.z -= pad;
        mx.x += pad; mx.y += pad; mx.z += pad;
        clusterLo[o] = make_float4(mn.x, mn.y, mn.z, __int_as_float(first));
        clusterHi[o] = make_float4(mx.x, mx.y, mx.z, __int_as_float(cnt));
        o++;
    }
}

__global__ void k_gather(const float* __restrict__ verts, const float* __restrict__ canon, const uint8_t* __restrict__ degenerate,
                         const int* __restrict__ vals, int n, float4* ta, float4* tb, float4* tc, float4* tn, float4* oa, float4* ob,
                         float4* oc) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int id = vals[i];
    const float* c = canon + (size_t)id * 9;
    const float* o = verts + (size_t)id * 9;
    float idf = __int_as_float(id);
    ta[i] = make_float4(c[0], c[1], c[2], idf);
    tb[i] = make_float4(c[3], c[4], c[5], degenerate[id] ? 1.f : 0.f);
    tc[i] = make_float4(c[6], c[7], c[8], 0.f);
    // geometric normal (B-A) x (C-A) in fp64, rounded once: the depth of a hit is (A-eye).Ng / D.Ng, and an fp32 cross product
    // of two nearly parallel edges (the long thin triangles of CAD meshes) would lose most of its digits
    double e1x = (double)c[3] - c[0], e1y = (double)c[4] - c[1], e1z = (double)c[5] - c[2];
    double e2x = (double)c[6] - c[0], e2y = (double)c[7] - c[1], e2z = (double)c[8] - c[2];
    tn[i] = make_float4((float)(e1y * e2z - e1z * e2y), (float)(e1z * e2x - e1x * e2z), (float)(e1x * e2y - e1y * e2x), 0.f);
    oa[i] = make_float4(o[0], o[1], o[2], idf);
    ob[i] = make_float4(o[3], o[4], o[5], 0.f);
    oc[i] = make_float4(o[6], o[7], o[8], 0.f);
}

template <class T>
T* dalloc(size_t n) {
    T* p = nullptr;
    IPP_CUDA_CHECK(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
    return p;
}

}  // namespace

void build_bvh(const float* d_verts, const float* d_canon, const float* d_centroid, const uint8_t* d_degenerate, int T,
               const float bbmin[3], const float bbmax[3], DeviceScene& sc, cudaStream_t st) {
    const int n = T;
    sc.T = T;
    sc.Tpad = (T + 31) / 32 * 32;
    for (int k = 0; k < 3; ++k) { sc.bbmin[k] = bbmin[k]; sc.bbmax[k] = bbmax[k]; }
    sc.ta = dalloc<float4>(n); sc.tb = dalloc<float4>(n); sc.tc = dalloc<float4>(n); sc.tn = dalloc<float4>(n);
    sc.oa = dalloc<float4>(n); sc.ob = dalloc<float4>(n); sc.oc = dalloc<float4>(n);
    sc.nNodes = std::max(n - 1, 1);
    sc.nodes = dalloc<BvhNode>(sc.nNodes);
    if (n == 0) return;

    unsigned long long *keys = dalloc<unsigned long long>(n), *keys2 = dalloc<unsigned long long>(n);
    int *vals = dalloc<int>(n), *vals2 = dalloc<int>(n);
    float ext[3];
    float maxExt = 0.f;
    for (int k = 0; k < 3; ++k) { ext[k] = bbmax[k] - bbmin[k]; maxExt = std::max(maxExt, ext[k]); }
    float3 lo = make_float3(bbmin[0], bbmin[1], bbmin[2]);
    float3 inv = make_float3(ext[0] > 0 ? 1.f / ext[0] : 0.f, ext[1] > 0 ? 1.f / ext[1] : 0.f, ext[2] > 0 ? 1.f / ext[2] : 0.f);
    const int B = 256, G = (n + B - 1) / B;
    k_morton<<<G, B, 0, st>>>(d_centroid, n, lo, inv, keys, vals);
    size_t tmpBytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmpBytes, keys, keys2, vals, vals2, n, 0, 63, st);
    void* tmp = nullptr;
    IPP_CUDA_CHECK(cudaMalloc(&tmp, std::max<size_t>(tmpBytes, 16)));
    cub::DeviceRadixSort::SortPairs(tmp, tmpBytes, keys, keys2, vals, vals2, n, 0, 63, st);

    k_gather<<<G, B, 0, st>>>(d_verts, d_canon, d_degenerate, vals2, n, sc.ta, sc.tb, sc.tc, sc.tn, sc.oa, sc.ob, sc.oc);

    float4 *leafMin = dalloc<float4>(n), *leafMax = dalloc<float4>(n), *nodeMin = dalloc<float4>(n), *nodeMax = dalloc<float4>(n);
    int *left = dalloc<int>(n), *right = dalloc<int>(n), *rlo = dalloc<int>(n), *rhi = dalloc<int>(n), *parentI = dalloc<int>(n),
        *parentL = dalloc<int>(n), *arrive = dalloc<int>(n);
    IPP_CUDA_CHECK(cudaMemsetAsync(arrive, 0, (size_t)n * sizeof(int), st));
    const float pad = 1e-6f * std::max(maxExt, 1.f);
    if (n >= 2) {
        k_karras<<<G, B, 0, st>>>(keys2, n, left, right, rlo, rhi, parentI, parentL);
        k_refit<<<G, B, 0, st>>>(d_verts, vals2, n, left, right, parentI, parentL, leafMin, leafMax, nodeMin, nodeMax, arrive);
        k_emit<<<G, B, 0, st>>>(n, left, right, rlo, rhi, leafMin, leafMax, nodeMin, nodeMax, pad, sc.nodes);
        // flat cluster table for the binned visibility kernel (arrive / parentI are free again: counts / offsets); clusters grow
        // until a camera pose can sort all of them (kMaxBinLeaves), up to kMaxClusterTris triangles each
        int *counts = arrive, *offsets = parentI;
        size_t scanBytes = scan_tmp_bytes(n);
        void* scanTmp = nullptr;
        IPP_CUDA_CHECK(cudaMalloc(&scanTmp, scanBytes));
        for (int C = kLeafMax;; C *= 2) {
            k_cluster_count<<<G, B, 0, st>>>(n, C, left, right, rlo, rhi, counts);
            IPP_CUDA_CHECK(cudaMemsetAsync(counts + (n - 1), 0, sizeof(int), st));
            exclusive_scan_i32(counts, offsets, n, scanTmp, scanBytes, st);
            IPP_CUDA_CHECK(cudaMemcpyAsync(&sc.nLeaves, offsets + (n - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
            IPP_CUDA_CHECK(cudaStreamSynchronize(st));
            sc.leafCap = C;
            if (sc.nLeaves <= kMaxBinLeaves || C * 2 > kMaxClusterTris) break;
        }
        sc.leafLo = dalloc<float4>(sc.nLeaves);
        sc.leafHi = dalloc<float4>(sc.nLeaves);
        k_cluster_fill<<<G, B, 0, st>>>(n, sc.leafCap, left, right, rlo, rhi, leafMin, leafMax, nodeMin, nodeMax, pad, counts, offsets, sc.leafLo,
                                        sc.leafHi);
        IPP_CUDA_CHECK(cudaStreamSynchronize(st));
        cudaFree(scanTmp);
    } else {
        // single triangle: root with one leaf child and one empty child
        BvhNode nd;
        nd.q0 = make_float4(bbmin[0] - pad, bbmin[1] - pad, bbmin[2] - pad, bbmax[0] + pad);
        nd.q1 = make_float4(bbmax[1] + pad, bbmax[2] + pad, 1e30f, 1e30f);
        nd.q2 = make_float4(1e30f, -1e30f, -1e30f, -1e30f);
        int l0 = ~((0 << kLeafShift) | 1), l1 = ~0;
        float f0, f1;
        memcpy(&f0, &l0, 4);
        memcpy(&f1, &l1, 4);
        nd.q3 = make_float4(f0, f1, 0.f, 0.f);
        IPP_CUDA_CHECK(cudaMemcpyAsync(sc.nodes, &nd, sizeof(nd), cudaMemcpyHostToDevice, st));
        sc.nLeaves = 1;
        sc.leafCap = kLeafMax;
        sc.leafLo = dalloc<float4>(1);
        sc.leafHi = dalloc<float4>(1);
        float4 lo4 = make_float4(nd.q0.x, nd.q0.y, nd.q0.z, 0.f), hi4 = make_float4(nd.q0.w, nd.q1.x, nd.q1.y, 0.f);
        const int first = 0, cnt = 1;
        memcpy(&lo4.w, &first, 4);
        memcpy(&hi4.w, &cnt, 4);
        IPP_CUDA_CHECK(cudaMemcpyAsync(sc.leafLo, &lo4, sizeof(lo4), cudaMemcpyHostToDevice, st));
        IPP_CUDA_CHECK(cudaMemcpyAsync(sc.leafHi, &hi4, sizeof(hi4), cudaMemcpyHostToDevice, st));
    }
    IPP_CUDA_CHECK(cudaStreamSynchronize(st));
    IPP_CUDA_CHECK(cudaGetLastError());
    cudaFree(keys); cudaFree(keys2); cudaFree(vals); cudaFree(vals2); cudaFree(tmp);
    cudaFree(leafMin); cudaFree(leafMax); cudaFree(nodeMin); cudaFree(nodeMax);
    cudaFree(left); cudaFree(right); cudaFree(rlo); cudaFree(rhi); cudaFree(parentI); cudaFree(parentL); cudaFree(arrive);
}

// Deterministic pseudo-random order of a set of keys: sort by key * 0x9E3779B1 mod 2^32 (a bijection of the 32-bit integers),
// then undo the multiplication.  Every rank that holds the same set gets the same order, and a contiguous chunk of the result is
// a random sample of the set (neighbouring keys -- edges that leave the same viewpoint and cost alike -- end up on different ranks).
namespace {
__global__ void k_mul_u32(const int* __restrict__ in, unsigned* __restrict__ out, int n, unsigned m) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (unsigned)in[i] * m;
}
__global__ void k_mul_back(unsigned* __restrict__ io, int n, unsigned m) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) io[i] = io[i] * m;
}
}  // namespace
size_t sort_tmp_bytes(int n) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, bytes, (const unsigned*)nullptr, (unsigned*)nullptr, n, 0, 32);
    return std::max<size_t>(bytes, 16) + (size_t)n * sizeof(unsigned) + 512;
}
void scrambled_order_i32(const int* d_in, int* d_out, int n, void* d_tmp, size_t tmpBytes, cudaStream_t st) {
    if (n <= 0) return;
    unsigned* h = (unsigned*)d_tmp;  // the hashed keys, then the CUB workspace
    char* work = (char*)d_tmp + (((size_t)n * sizeof(unsigned) + 255) / 256) * 256;
    size_t workBytes = tmpBytes - (size_t)(work - (char*)d_tmp);
    k_mul_u32<<<(n + 255) / 256, 256, 0, st>>>(d_in, h, n, 0x9E3779B1u);
    cub::DeviceRadixSort::SortKeys(work, workBytes, h, (unsigned*)d_out, n, 0, 32, st);
    k_mul_back<<<(n + 255) / 256, 256, 0, st>>>((unsigned*)d_out, n, 0x0E8B2F51u);
}
size_t scan_tmp_bytes(int n) {
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, (const int*)nullptr, (int*)nullptr, n);
    return std::max<size_t>(bytes, 16);
}
void exclusive_scan_i32(const int* d_in, int* d_out, int n, void* d_tmp, size_t tmpBytes, cudaStream_t st) {
    cub::DeviceScan::ExclusiveSum(d_tmp, tmpBytes, d_in, d_out, n, st);
}

void free_bvh(DeviceScene& sc) {
    cudaFree(sc.ta); cudaFree(sc.tb); cudaFree(sc.tc); cudaFree(sc.tn);
    cudaFree(sc.oa); cudaFree(sc.ob); cudaFree(sc.oc);
    cudaFree(sc.nodes);
    cudaFree(sc.leafLo); cudaFree(sc.leafHi);
    sc.leafLo = sc.leafHi = nullptr;
    sc.nLeaves = 0;
    sc.ta = sc.tb = sc.tc = sc.tn = sc.oa = sc.ob = sc.oc = nullptr;
    sc.nodes = nullptr;
}

}  // namespace ipp
