// host_setup.h -- the once-per-problem pixel layout of the fp32 kernel, plain C++ (no CUDA) so that
// the C-ABI (mcalf_api.cu) and the host emulation used by the CPU tests build the same tables.
#pragma once
#include <math.h>

#include <algorithm>
#include <vector>

namespace mcalf {

// <= 256 consecutive pixels of the concatenated fit-window array sharing one fp64 reference rho_s
struct ChunkDesc {
    int start, len;
    float ds, inv_ds;     // max |delta| over the chunk (slightly widened), delta = rho - rho_s, and 1/ds
    double rho_s;
};

constexpr int CHUNK_PIXELS = 256;
constexpr double CHUNK_RHO_SPAN = 1.0 / 512.0;

// Chunks: consecutive pixels whose rho = lam_ref/lambda spans at most CHUNK_RHO_SPAN (so that the fp32
// offset delta = rho - rho_s keeps ~2^-34 absolute accuracy; fit-window gaps and non-uniform grids
// simply start a new chunk), and delta as a two-float per pixel.
inline void build_chunks(const double *wave, int npix, double lam_ref, std::vector<ChunkDesc> &chunks,
                         std::vector<float> &dhi, std::vector<float> &dlo) {
    chunks.clear();
    dhi.assign(npix, 0.0f);
    dlo.assign(npix, 0.0f);
    int start = 0;
    while (start < npix) {
        double rmin = lam_ref / wave[start], rmax = rmin;
        int len = 1;
        while (start + len < npix && len < CHUNK_PIXELS) {
            const double r = lam_ref / wave[start + len];
            const double nmin = std::min(rmin, r), nmx = std::max(rmax, r);
            if (nmx - nmin > CHUNK_RHO_SPAN) break;
            rmin = nmin;
            rmax = nmx;
            ++len;
        }
        ChunkDesc cd;
        cd.start = start;
        cd.len = len;
        cd.rho_s = 0.5 * (rmin + rmax);
        float dmin = 3e38f, dmax = -3e38f;
        for (int i = start; i < start + len; ++i) {
            const double d = lam_ref / wave[i] - cd.rho_s;
            dhi[i] = (float)d;
            dlo[i] = (float)(d - (double)dhi[i]);
            dmin = std::min(dmin, dhi[i]);
            dmax = std::max(dmax, dhi[i]);
        }
        // widened a little so the classification bound also covers the dropped low part
        cd.ds = std::max(fabsf(dmin), fabsf(dmax)) * (1.0f + 1e-6f) + 1e-12f;
        cd.inv_ds = 1.0f / cd.ds;
        chunks.push_back(cd);
        start += len;
    }
}

// Pair tables of the fp32 kernel: a lane holds the pixels k and k + 32 of a 64-pixel row pair in one
// register pair, so the tables store {v[i], v[i + 32]} at i (zero where i + 32 leaves the chunk).  Padded
// by a full chunk: the kernel loads all 256 slots of a chunk without bounds tests (slots beyond a short
// chunk read the next chunk's entries or the zero padding; nothing computed from them is stored).
struct PairF { float x, y; };
inline void build_pair_table(const std::vector<ChunkDesc> &chunks, const std::vector<float> &v, std::vector<PairF> &out) {
    out.assign(v.size() + CHUNK_PIXELS, PairF{0.0f, 0.0f});
    for (const ChunkDesc &cd : chunks)
        for (int i = cd.start; i < cd.start + cd.len; ++i) {
            out[i].x = v[i];
            out[i].y = (i + 32 < cd.start + cd.len) ? v[i + 32] : 0.0f;
        }
}

}  // namespace mcalf
