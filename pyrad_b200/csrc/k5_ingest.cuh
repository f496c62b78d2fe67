// k5_ingest.cuh -- HITRAN-online CSV text -> SoA line arrays, on the device.
//
// Replaces readHitranOnlineFile / gatherData (pyradUtilities.py:421-448, 173-189) and the per-line object build
// (pyradClasses.py:350-359) for the engine's purposes: rows
//     molec_id,local_iso_id,nu,sw,a,elower,gamma_air,gamma_self,delta_air,n_air
// are parsed with the exact decimal -> binary64 conversion of numparse.cuh (what float(cell) returns), kept when
// waveMin < nu < waveMax (strict, :437-438), duplicate wavenumbers collapse with the LAST row winning (the
// reference keys a dict by nu, :447), and the survivors land in the engine's SoA columns in file order -- the raw
// text is the only thing that crosses PCIe.
//
// Byte work, HBM bound: one pass marks the newlines (coalesced 16-byte loads), CUB compacts their positions and
// later the kept-row flags (plumbing), one thread per row parses its ~100 bytes.
#pragma once
#include <cub/cub.cuh>
#include "common.cuh"
#include "numparse.cuh"

namespace prb {

constexpr unsigned char ROW_SKIP = 0, ROW_KEEP = 1, ROW_OUT = 2, ROW_BAD = 3;

struct IngestCols {
    double *nu, *sw, *a, *elower, *gair, *gself, *delta, *nair;
};

__global__ void __launch_bounds__(256)
k5_mark_newlines(const char *__restrict__ text, int64_t n, unsigned char *__restrict__ is_nl,
                 unsigned long long *__restrict__ n_newlines) {
    // 16 bytes per thread (the text buffer is padded to a multiple of 16); the newline count is a plain integer
    // count (one atomic per warp), needed to size the row tables
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t i0 = v * 16;
    uint4 w = make_uint4(0, 0, 0, 0);
    if (i0 < n) w = *reinterpret_cast<const uint4 *>(text + i0);
    const unsigned int ws[4] = {w.x, w.y, w.z, w.w};
    unsigned int out[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        unsigned int o = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const bool nl = ((ws[k] >> (8 * b)) & 0xffu) == (unsigned int)'\n' && (i0 + 4 * k + b) < n;
            o |= (nl ? 1u : 0u) << (8 * b);
        }
        out[k] = o;
    }
    if (i0 < n) *reinterpret_cast<uint4 *>(is_nl + i0) = make_uint4(out[0], out[1], out[2], out[3]);
    int cnt = __popc(out[0]) + __popc(out[1]) + __popc(out[2]) + __popc(out[3]);
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(n_newlines, (unsigned long long)cnt);
}

// Row r spans [start_r, end_r): start_0 = 0, start_r = nl_pos[r-1] + 1; end_r = nl_pos[r] (or n for a last row
// without a trailing newline).
constexpr int K5_ROWS = 128;            // rows (threads) per block
constexpr int K5_STAGE = 24 * 1024;     // shared-memory staging of the block's rows (they are contiguous in the text)

__global__ void __launch_bounds__(K5_ROWS)
k5_parse_rows(const char *__restrict__ text, int64_t n, const int64_t *__restrict__ nl_pos, int64_t n_nl, int64_t n_rows,
              double wave_min, double wave_max, IngestCols tmp, unsigned char *__restrict__ state,
              unsigned long long *__restrict__ first_bad) {
    __shared__ __align__(16) char stage[K5_STAGE];
    // the block's rows [r0, r1) are one contiguous byte range: pull it into shared memory with coalesced 16-byte
    // loads, then every thread walks its own row there (byte-wise walks of global memory do not coalesce)
    const int64_t r0 = (int64_t)blockIdx.x * K5_ROWS;
    const int64_t r1 = min(r0 + (int64_t)K5_ROWS, n_rows);
    const int64_t blk_b = r0 == 0 ? 0 : nl_pos[r0 - 1] + 1;
    const int64_t blk_e = (r1 - 1) < n_nl ? nl_pos[r1 - 1] : n;
    const int64_t a0 = blk_b & ~int64_t(15);
    const bool staged = blk_e - a0 <= K5_STAGE;
    if (staged) {
        const int64_t span = (blk_e - a0 + 15) & ~int64_t(15);              // the text buffer is padded to 16 bytes
        for (int64_t i = (int64_t)threadIdx.x * 16; i < span; i += (int64_t)K5_ROWS * 16)
            *reinterpret_cast<uint4 *>(stage + i) = *reinterpret_cast<const uint4 *>(text + a0 + i);
    }
    __syncthreads();
    const int64_t r = r0 + threadIdx.x;
    if (r >= n_rows) return;
    const int64_t b = r == 0 ? 0 : nl_pos[r - 1] + 1;
    const int64_t e = r < n_nl ? nl_pos[r] : n;
    const char *base = staged ? stage - a0 : text;
    const char *p = base + b, *end = base + e;
    unsigned char st = ROW_BAD;
    if (p < end && *p == '#') {
        st = ROW_SKIP;                                            // comment row
    } else if (p == end && r == n_rows - 1) {
        st = ROW_SKIP;                                            // nothing after the final newline
    } else {
        // split at commas: fields 2..9 are the numbers the reference reads
        const char *fb[10], *fe[10];
        int nf = 0;
        const char *q = p;
        fb[0] = p;
        for (; q < end; ++q) {
            if (*q == ',') {
                fe[nf] = q;
                if (++nf == 10) break;
                fb[nf] = q + 1;
            }
        }
        if (nf < 10) { fe[nf] = end; ++nf; }
        double nu;
        if (nf >= 10 && parse_double(fb[2], fe[2], &nu)) {
            if (wave_min < nu && nu < wave_max) {
                double v[7];
                bool ok = true;
#pragma unroll
                for (int k = 0; k < 7; ++k) ok = parse_double(fb[3 + k], fe[3 + k], &v[k]) && ok;
                if (ok) {
                    st = ROW_KEEP;
                    tmp.nu[r] = nu; tmp.sw[r] = v[0]; tmp.a[r] = v[1]; tmp.elower[r] = v[2];
                    tmp.gair[r] = v[3]; tmp.gself[r] = v[4]; tmp.delta[r] = v[5]; tmp.nair[r] = v[6];
                }
            } else {
                st = ROW_OUT;
                tmp.nu[r] = nu;
            }
        }
    }
    state[r] = st;
    if (st == ROW_BAD) atomicMin(first_bad, (unsigned long long)r);
}

// xsc cross-section tables (returnXscFileContents pyradUtilities.py:680-696): a row is kept when, stripped and split at
// runs of SPACES, it has exactly two tokens that both parse as numbers; anything else is skipped (the reference logs
// "line skipped").  Leading '#' rows are comments (openReturnLines :100-101).
__global__ void __launch_bounds__(K5_ROWS)
k5_parse_xsc_rows(const char *__restrict__ text, int64_t n, const int64_t *__restrict__ nl_pos, int64_t n_nl, int64_t n_rows,
                  double *__restrict__ wn, double *__restrict__ xs, int32_t *__restrict__ keep) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const int64_t b = r == 0 ? 0 : nl_pos[r - 1] + 1;
    const int64_t e = r < n_nl ? nl_pos[r] : n;
    const char *p = text + b, *end = text + e;
    while (p < end && is_ws(*p)) ++p;                             // line.strip()
    while (end > p && is_ws(end[-1])) --end;
    int32_t k = 0;
    if (p < end && *p != '#') {
        const char *t1 = p;
        while (t1 < end && *t1 != ' ') ++t1;                      // first token: up to the first space
        const char *t2 = t1;
        while (t2 < end && *t2 == ' ') ++t2;                      // the run of spaces
        const char *t3 = t2;
        while (t3 < end && *t3 != ' ') ++t3;                      // second token must reach the end of the row
        double a, c;
        if (t1 > p && t2 > t1 && t3 == end && t3 > t2 && parse_double(p, t1, &a) && parse_double(t2, t3, &c)) {
            wn[r] = a;
            xs[r] = c;
            k = 1;
        }
    }
    keep[r] = k;
}

__global__ void __launch_bounds__(256)
k5_scatter2(const double *__restrict__ a, const double *__restrict__ b, const int32_t *__restrict__ keep,
            const int32_t *__restrict__ pos, int64_t n_rows, double *__restrict__ oa, double *__restrict__ ob) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows || !keep[r]) return;
    oa[pos[r]] = a[r];
    ob[pos[r]] = b[r];
}

// The reference's dict keyed by nu: a later row with the same wavenumber replaces the earlier one.  Files are
// ascending in nu, so duplicates are adjacent (only comment rows can sit between them); whatever this fast path misses
// shows up as two equal neighbours among the kept rows and takes the sort path (k5_finalize flags it).
__global__ void __launch_bounds__(256)
k5_resolve_duplicates(const double *__restrict__ nu, const unsigned char *__restrict__ state, int64_t n_rows,
                      int32_t *__restrict__ keep) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    int32_t k = state[r] == ROW_KEEP;
    if (k) {
        int64_t j = r + 1;
        while (j < n_rows && state[j] == ROW_SKIP) ++j;
        if (j < n_rows && state[j] == ROW_KEEP && nu[j] == nu[r]) k = 0;
    }
    keep[r] = k;
}

__global__ void __launch_bounds__(256)
k5_scatter_kept(IngestCols tmp, const int32_t *__restrict__ keep, const int32_t *__restrict__ pos, int64_t n_rows,
                IngestCols dst) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows || !keep[r]) return;
    const int64_t o = pos[r];
    dst.nu[o] = tmp.nu[r]; dst.sw[o] = tmp.sw[r]; dst.a[o] = tmp.a[r]; dst.elower[o] = tmp.elower[r];
    dst.gair[o] = tmp.gair[r]; dst.gself[o] = tmp.gself[r]; dst.delta[o] = tmp.delta[r]; dst.nair[o] = tmp.nair[r];
}

// Out-of-order segment files (the reference's reader keys a dict by nu and takes any order, pyradUtilities.py:421-448):
// the kept rows are stably sorted by wavenumber (row numbers as the payload of a radix sort), which also brings equal
// wavenumbers from anywhere in the file next to each other in FILE order -- the last of every run is the row the
// reference's dict ends up with.
__global__ void __launch_bounds__(256) k5_iota(int32_t *__restrict__ v, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = (int32_t)i;
}
__global__ void __launch_bounds__(256)
k5_keep_last_of_run(const double *__restrict__ nu_sorted, int64_t n, int32_t *__restrict__ keep) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keep[i] = (i + 1 == n || nu_sorted[i + 1] != nu_sorted[i]) ? 1 : 0;
}
// dst[pos[i]] = src[perm[i]] for the kept entries of the sorted order
__global__ void __launch_bounds__(256)
k5_gather_sorted(IngestCols src, const int32_t *__restrict__ perm, const int32_t *__restrict__ keep,
                 const int32_t *__restrict__ pos, int64_t n, IngestCols dst) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !keep[i]) return;
    const int64_t r = perm[i], o = pos[i];
    dst.nu[o] = src.nu[r]; dst.sw[o] = src.sw[r]; dst.a[o] = src.a[r]; dst.elower[o] = src.elower[r];
    dst.gair[o] = src.gair[r]; dst.gself[o] = src.gself[r]; dst.delta[o] = src.delta[r]; dst.nair[o] = src.nair[r];
}

// What prb_upload_lines checks for a caller-supplied list, for the ingested one: strictly ascending nu0 and max |S296|.
__global__ void __launch_bounds__(256)
k5_finalize(const double *__restrict__ nu, const double *__restrict__ sw, int64_t n, unsigned long long *__restrict__ smax_bits,
            unsigned int *__restrict__ unsorted) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long b = 0;
    unsigned int bad = 0;
    if (i < n) {
        b = (unsigned long long)__double_as_longlong(fabs(sw[i]));                // |S| >= 0: bit order == value order
        // not STRICTLY ascending: out of order, or two kept rows of one wavenumber that k5_resolve_duplicates did not see
        // as neighbours (rows outside the wavenumber range stood between them) -> the sort path keeps the last of the run
        if (i + 1 < n && !(nu[i + 1] > nu[i])) bad = 1;
    }
    bad = __reduce_or_sync(0xffffffffu, bad);
    // warp max of a 64-bit key: two 32-bit reductions (high word first)
    const unsigned int hi = __reduce_max_sync(0xffffffffu, (unsigned int)(b >> 32));
    const unsigned int lo = __reduce_max_sync(0xffffffffu, (unsigned int)(b >> 32) == hi ? (unsigned int)b : 0u);
    if ((threadIdx.x & 31) == 0) {
        atomicMax(smax_bits, ((unsigned long long)hi << 32) | lo);
        if (bad) atomicOr(unsorted, 1u);
    }
}

}  // namespace prb
