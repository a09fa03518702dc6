// TEST-ONLY: runs the __host__ __device__ per-lane logic of the kernels (Huffman LUTs, BitReader,
// decode_block, idct_put_block, fdct_block, quant_*) on the CPU, one emulated lane at a time, so
// the intricate parts can be checked against the oracle in a container without a GPU.  This is
// not a product path: nothing outside tests/ builds or loads it, and libamvcuda never links it.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../amv-codec-tools_b200/csrc/amv_common.cuh"
#include "../../amv-codec-tools_b200/csrc/amv_tables.cuh"
#include "../../amv-codec-tools_b200/csrc/amv_dct.cuh"
#include "../../amv-codec-tools_b200/csrc/amv_vlc.cuh"
#include "../../amv-codec-tools_b200/csrc/amv_kernels.h"

using namespace amv;

static VlcTables g_vlc;
static DequantTables g_dq;
static EncHuffTables g_eh;
static bool g_init;
static void init() {
    if (g_init) return;
    build_vlc_tables(g_vlc); build_dequant_tables(g_dq); build_enc_huff_tables(g_eh);
    g_init = true;
}

// same rule as k_unstuff, sequential
static std::vector<uint8_t> unstuff(const uint8_t *pkt, uint32_t size, int *st) {
    std::vector<uint8_t> v;
    const uint32_t npay = size >= 4 ? size - 4 : 0;
    std::vector<uint8_t> in(pkt + (size >= 4 ? 2 : 0), pkt + (size >= 4 ? 2 : 0) + npay);
    in.push_back(0xff); in.push_back(0xd9);
    uint32_t prev = 0;
    for (size_t i = 0; i < in.size(); i++) {
        const uint32_t x = in[i];
        const bool after_ff = prev == 0xff && i > 0;
        const bool drop = after_ff && (x == 0 || x == 0xff);
        const bool term = after_ff && !(x == 0 || x == 0xff || (x >= 0xd0 && x <= 0xd7));
        if (term) { if (i != in.size() - 1) *st |= AMV_ST_MARKER; break; }
        if (!drop) v.push_back((uint8_t)x);
        prev = x;
    }
    return v;
}

struct CountSink { int dcv; void dc(int d) { dcv = d; } void ac(int, int) {} };
// producer side of the token format: (column byte offset << 16) | int16 value, like k_vlc_tokens
struct TokSink {
    std::vector<uint32_t> *tok; const uint32_t *tz; int dcv; uint32_t nac;
    void dc(int d) { dcv = d; }
    void ac(int k, int level) {
        const uint32_t z = tz[k];
        tok->push_back((z & 0xffff0000u) | ((uint32_t)(level * (int)(z & 0xffffu)) & 0xffffu));
        nac++;
    }
};

static void walk(const uint32_t *words, uint32_t nwords, uint32_t sbit, uint32_t sph, uint32_t end_bit, LaneExit &ex) {
    BitReader br; br.init(words, nwords, sbit);
    uint32_t phase = sph, nb = 0; int dc[3] = { 0, 0, 0 };
    CountSink sink;
    while (br.bitpos() < end_bit) {
        walk_block(br, g_vlc.e, g_vlc.base, phase >= 4, sink);
        dc[phase < 4 ? 0 : phase - 3] += sink.dcv;
        phase = phase == 5 ? 0 : phase + 1; nb++;
    }
    ex.bitpos = br.bitpos(); ex.phase = phase; ex.nblocks = nb; ex.dc[0] = dc[0]; ex.dc[1] = dc[1]; ex.dc[2] = dc[2];
}

extern "C" {

int emul_vlc_entries(void) { init(); return g_vlc.count; }

// decode one frame with P = 1 << log2p emulated lanes; returns status, *rounds = sync rounds used
int emul_decode_frame(const uint8_t *pkt, uint32_t size, int w, int h, uint8_t *py, uint8_t *pu, uint8_t *pv,
                      int log2p, int *rounds) {
    init();
    int st = 0;
    std::vector<uint8_t> scan = unstuff(pkt, size, &st);
    const uint32_t U = (uint32_t)scan.size();
    scan.resize(((U + 15) & ~15u) + 32, 0);
    const uint32_t *words = reinterpret_cast<const uint32_t *>(scan.data());
    const uint32_t nwords = (U + 3) >> 2, total_bits = U * 8;
    const Geom g = make_geom(w, h);
    const int P = 1 << log2p;
    std::vector<LaneStart> starts(P);
    if (rounds) *rounds = 0;
    if (log2p == 0) { starts[0] = { 0, 0, (uint32_t)g.nblk, { 1024, 1024, 1024 } }; }
    else {
        const uint32_t L = (((total_bits + P - 1) >> log2p) + 31u) & ~31u;
        std::vector<uint32_t> sb(P), sp(P, 0), eb(P);
        std::vector<LaneExit> ex(P);
        for (int p = 0; p < P; p++) {
            uint64_t s = (uint64_t)p * L, e = (uint64_t)(p + 1) * L;
            sb[p] = (uint32_t)(s < total_bits ? s : total_bits); eb[p] = (uint32_t)(e < total_bits ? e : total_bits);
            walk(words, nwords, sb[p], 0, eb[p], ex[p]);
        }
        int r = 1;
        for (int it = 0; it < P; it++) {
            bool any = false;
            std::vector<LaneExit> old = ex;
            for (int p = 0; p < P; p++) {
                uint32_t nb = p ? old[p - 1].bitpos : 0, np = p ? old[p - 1].phase : 0;
                if (nb != sb[p] || np != sp[p]) { any = true; sb[p] = nb; sp[p] = np; walk(words, nwords, nb, np, eb[p], ex[p]); }
            }
            if (!any) break;
            r++;
        }
        if (rounds) *rounds = r;
        uint32_t first = 0; int d[3] = { 0, 0, 0 };
        const int q0l = g_dq.zq[0][0] >> 8, q0c = g_dq.zq[1][0] >> 8;
        for (int p = 0; p < P; p++) {
            starts[p].bitpos = sb[p]; starts[p].first_block = first; starts[p].nblocks = ex[p].nblocks;
            starts[p].pred[0] = 1024 + q0l * d[0]; starts[p].pred[1] = 1024 + q0c * d[1]; starts[p].pred[2] = 1024 + q0c * d[2];
            first += ex[p].nblocks; d[0] += ex[p].dc[0]; d[1] += ex[p].dc[1]; d[2] += ex[p].dc[2];
        }
    }
    // pass 1 (k_vlc_tokens): tokens + per-block (count, offset); pass 2 (k_idct): tokens -> pixels
    std::vector<uint32_t> tok;
    std::vector<uint32_t> boff(g.nblk, 0);
    std::vector<uint8_t> have(g.nblk, 0);
    for (int p = 0; p < P; p++) {
        uint32_t first = starts[p].first_block, count = starts[p].nblocks;
        if (log2p) {
            if (first >= (uint32_t)g.nblk) count = 0; else if (first + count > (uint32_t)g.nblk) count = g.nblk - first;
            if (p == P - 1 && first + starts[p].nblocks < (uint32_t)g.nblk) { count = g.nblk - first; st |= AMV_ST_OVERRUN; }
        }
        BitReader br; br.init(words, nwords, starts[p].bitpos);
        int pred[3] = { starts[p].pred[0], starts[p].pred[1], starts[p].pred[2] };
        int b = first % 6;
        for (uint32_t i = 0; i < count; i++) {
            const int tq = b >= 4, comp = b < 4 ? 0 : b - 3;
            const uint32_t start = (uint32_t)tok.size();
            have[first + i] = 1;
            tok.push_back(0);
            TokSink sink; sink.tok = &tok; sink.tz = g_dq.tz[tq]; sink.nac = 0;
            st |= walk_block(br, g_vlc.e, g_vlc.base, tq, sink);
            pred[comp] += sink.dcv * (int)(g_dq.zq[tq][0] >> 8);
            tok[start] = (uint32_t)pred[comp] & 0xffffu;
            boff[first + i] = (sink.nac << kTokCountShift) | start;
            if (++b == 6) b = 0;
        }
        if (count && br.bitpos() > U * 8u) st |= AMV_ST_OVERRUN;
    }
    tok.resize(tok.size() + 8, 0);
    for (int blk = 0; blk < g.nblk; blk++) {
        if (!have[blk]) continue;
        const int mb = blk / 6, b = blk % 6, mx = mb % g.mbw, my = mb / g.mbw;
        const int comp = b < 4 ? 0 : b - 3;
        int16_t coef[64] = { 0 };
        const uint32_t *tp = tok.data() + (boff[blk] & ((1u << kTokCountShift) - 1u));
        const uint32_t nac = boff[blk] >> kTokCountShift;
        for (uint32_t a = 0; a <= nac; a++) {
            const uint32_t t = tp[a], o = t >> 16;                      // column byte offset -> raster index
            coef[(o / 128) * 2 + ((o % 128) >> 1)] = (int16_t)(t & 0xffffu);
        }
        uint32_t c[32], o[16];
        for (int k = 0; k < 32; k++) c[k] = (uint16_t)coef[2 * k] | ((uint32_t)(uint16_t)coef[2 * k + 1] << 16);
        idct_put_block(c, o);
        uint8_t *pl = comp == 0 ? py : (comp == 1 ? pu : pv);
        const int ls = comp ? g.cw : g.w, vw = comp ? g.cw : g.w, vh = comp ? g.ch : g.h, r0 = comp ? g.c0 : g.y0;
        const int bx = comp ? mx * 8 : mx * 16 + (b & 1) * 8, by = comp ? my * 8 : my * 16 + (b >> 1) * 8;
        for (int yy = 0; yy < 8; yy++) {
            const int row = r0 - (by + yy);
            if (row < 0 || row >= vh) continue;
            for (int xx = 0; xx < 8; xx++)
                if (bx + xx < vw) pl[row * ls + bx + xx] = (uint8_t)(o[2 * yy + (xx >> 2)] >> (8 * (xx & 3)));
        }
    }
    return st;
}

void emul_idct(const int16_t *blocks, int n, uint8_t *out) {
    for (int i = 0; i < n; i++) {
        uint32_t c[32], o[16];
        for (int k = 0; k < 32; k++) c[k] = (uint16_t)blocks[64 * i + 2 * k] | ((uint32_t)(uint16_t)blocks[64 * i + 2 * k + 1] << 16);
        uint32_t lower = 0;                              // the kernel's choice: nothing below the second row -> short transform
        for (int k = 8; k < 32; k++) lower |= c[k];
        if (lower == 0) idct_put_block<2>(c, o); else idct_put_block<8>(c, o);
        memcpy(out + 64 * i, o, 64);
    }
}

// pixels (0..255 as int16) -> fdct -> quantised coefficients in raster order, like the kernel's stage A
void emul_fdct_quant(const int16_t *blocks, int n, int qscale, int16_t *out, int16_t *fdct_out) {
    uint32_t qm10[64];
    for (int t = 0; t < 64; t++) {
        int m = 8;
        if (t) { m = (kEncIntraBase[t] * qscale) >> 3; m = m < 1 ? 1 : (m > 255 ? 255 : m); }
        qm10[t] = ((1u << 22) / (uint32_t)(8 * m)) << 10;
    }
    for (int i = 0; i < n; i++) {
        int v[64];
        for (int k = 0; k < 64; k++) v[k] = blocks[64 * i + k];
        fdct_block(v);
        if (fdct_out) for (int k = 0; k < 64; k++) fdct_out[64 * i + k] = (int16_t)v[k];
        out[64 * i] = (int16_t)quant_dc(v[0]);
        for (int k = 1; k < 64; k++) out[64 * i + k] = (int16_t)quant_ac(v[k], qm10[k]);
    }
}

// div_by_magic over a range of dividends: returns the number of mismatches against the C division
int emul_div_magic_check(uint32_t d, uint32_t x0, uint32_t count, uint32_t step) {
    const uint32_t m = div_magic(d);
    int bad = 0;
    uint32_t x = x0;
    for (uint32_t i = 0; i < count; i++, x += step) {
        uint32_t rem;
        const uint32_t q = div_by_magic(x, d, m, rem);
        if (q != x / d || rem != x % d) bad++;
    }
    return bad;
}

// the encoder's transform in its regrouped forms (fdct_block_px: dot-product rows on packed bytes, written-out columns)
}  // extern "C"
template <int FORM>
static void fdct_form_run(const int16_t *blocks, int n, int16_t *fdct_out) {
    for (int i = 0; i < n; i++) {
        uint32_t px[16];
        for (int k = 0; k < 16; k++) {
            px[k] = 0;
            for (int j = 0; j < 4; j++) px[k] |= (uint32_t)(blocks[64 * i + 4 * k + j] & 0xff) << (8 * j);
        }
        int v[64];
        fdct_block_px<FORM>(px, v);
        for (int k = 0; k < 64; k++) fdct_out[64 * i + k] = (int16_t)v[k];
    }
}
extern "C" {
int emul_fdct_form(const int16_t *blocks, int n, int form, int16_t *fdct_out) {
    switch (form) {
    case 0: fdct_form_run<0>(blocks, n, fdct_out); return 0;
    case 1: fdct_form_run<1>(blocks, n, fdct_out); return 0;
    case 2: fdct_form_run<2>(blocks, n, fdct_out); return 0;
    case 3: fdct_form_run<3>(blocks, n, fdct_out); return 0;
    case 6: fdct_form_run<6>(blocks, n, fdct_out); return 0;
    case 7: fdct_form_run<7>(blocks, n, fdct_out); return 0;
    }
    return -1;
}

// worst-case magnitudes inside fdct_block for a block: reports whether every intermediate the
// 32-bit kernel shifts fits (|x| < 2^31) by redoing the pass in 64 bit
int emul_enc_huff(int idx) { init(); return (int)g_eh.e[idx]; }

}
