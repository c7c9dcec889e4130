// hamming_experiments.cuh -- round-2 schedules of the tensor-core Hamming scan that were measured and NOT adopted.
// Compiled only with -DUCFP_HAMMING_EXPERIMENTS (python -m ucfp_b200.build --experiments); selected at run time with
// UCFP_HAMMING_MMA_V=2|3|4 and UCFP_HAMMING_EPI_WARPS=8|16.  All of them pass the tensor-path parity tests; none beats the
// first-generation kernel (hamming_mma_scan_kernel), which therefore stays the product path.  Measured on one B200,
// 1024 queries x 250 M codes, whole scan (profiles/r02_hamming_schedules.md):
//     first generation (product)                                   10.2 ms
//     v2  two half-loads per item, one in flight      16 / 8 warps  12.9 / 12.3 ms
//     v3  next item's load in flight                  16 / 8 warps  13.9 (spills at 96 registers) / 13.4 ms
//     v4  four 128-column TMEM stages, N = 128 MMAs   16 / 8 warps  15.2 / 14.2 ms
// What the experiments established (scripts/micro/tmem_port.cu, ncu captures): the MMA stream alone runs at exactly 256 clk per
// 128 x 256 x 64 tile and tcgen05.ld at ~135-145 clk per tile even concurrently -- TMEM bandwidth is not the bound; the
// first generation spends ~420 clk per tile on the ALU pipe (VIMNMX3.S16x2 issues every 2.15 clk) plus ~90 clk of mbarrier
// round trip and ~150 clk of tcgen05.ld latency that all sixteen epilogue warps sit out together.  Hiding those latencies by
// double-buffering costs more than it gains: fewer or wider epilogue warps lose the cross-warp overlap, 16 warps with two
// register images do not fit 96 registers, a setmaxnreg variant hung on the device, and with N = 128 the single MMA-issuing
// warp (one mbarrier round trip per item) becomes the bottleneck.  Included inside namespace ucfp { namespace { } } by hamming.cu.
#pragma once

// ---- second generation of the stage-image scan -----------------------------------------------------------------------
// Same arithmetic, operand layout and shared-memory layout as hamming_mma_scan_kernel<true>; what changes is the schedule
// of the epilogue.  Round 1's kernel sat at 656 clk per 128 x 512-pair accumulator tile (tensor pipe 39 % active): all
// sixteen epilogue warps wait on the same mbarrier, issue their tcgen05.ld together and then stall on it -- the TMEM read
// port delivers ~64 B/clk per sub-partition, so the last warp of a sub-partition gets its columns ~256 clk after the first
// and its min/max work runs with nothing left to overlap.  Here every epilogue warp reads its columns as two halves and
// always has ONE half in flight while it reduces the other: the port stays busy across tile boundaries and the ALU work
// hides under it.  The TMEM stage is handed back as soon as the second half has landed.  kEpiW = 16: 64 columns per warp,
// halves of 32 columns (tcgen05.ld x16); kEpiW = 8: 128 columns per warp, halves of 64 columns (x32), half the per-item
// bookkeeping per column.  Warps: MMA issuer, one TMA thread, kEpiW epilogue warps.
__device__ __forceinline__ void tmem_ld_pack16_async(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.pack::16b.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_pack16_async(uint32_t taddr, uint32_t (&v)[32]) { tmem_ld64_pack16_async(taddr, v); }
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&v)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]),
                   "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]) :: "memory");
}
// 32 lanes x 128 columns, packed: 64 registers (one load per item for the 8-warp form of the third schedule)
__device__ __forceinline__ void tmem_ld_pack16_async(uint32_t taddr, uint32_t (&v)[64]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.pack::16b.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
                 "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]),
                   "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),
                   "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), "=r"(v[32]), "=r"(v[33]),
                   "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]),
                   "=r"(v[45]), "=r"(v[46]), "=r"(v[47]), "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]),
                   "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&v)[64]) {
    // the registers are tied to the statement in two halves (an asm statement takes at most 30 operands of this kind comfortably)
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]),
                   "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]),
                   "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31]) :: "memory");
    asm volatile("" : "+r"(v[32]), "+r"(v[33]), "+r"(v[34]), "+r"(v[35]), "+r"(v[36]), "+r"(v[37]), "+r"(v[38]), "+r"(v[39]), "+r"(v[40]), "+r"(v[41]), "+r"(v[42]),
                      "+r"(v[43]), "+r"(v[44]), "+r"(v[45]), "+r"(v[46]), "+r"(v[47]), "+r"(v[48]), "+r"(v[49]), "+r"(v[50]), "+r"(v[51]), "+r"(v[52]), "+r"(v[53]),
                      "+r"(v[54]), "+r"(v[55]), "+r"(v[56]), "+r"(v[57]), "+r"(v[58]), "+r"(v[59]), "+r"(v[60]), "+r"(v[61]), "+r"(v[62]), "+r"(v[63]) :: "memory");
}
// per-halfword signed max of D and min of D << 9 against the query's two bounds (see hamming_mma_scan_kernel)
template <int NREG>
__device__ __forceinline__ bool hamming_mma_hot_test(const uint32_t (&p)[NREG], uint32_t hi_pk, uint32_t lo_pk) {
    // Padding queries skip the reduction.  The branch also pins the schedule: ptxas keeps the tcgen05.ld issued just before
    // this test AHEAD of the min/max work (without a block boundary it sinks the load to the end of the reduction, reusing
    // the load's destination registers as temporaries, and nothing overlaps).
    if (hi_pk == kMmaNeverHiPk) return false;
    if (NREG >= 64) {   // few warps per sub-partition: two independent chains per stream keep the ALU pipe fed
        uint32_t mx0 = hi_pk, mx1 = hi_pk, mn0 = lo_pk, mn1 = lo_pk;
#pragma unroll
        for (int c = 0; c < NREG; c += 4) {
            mx0 = __vimax3_s16x2(mx0, p[c], p[c + 1]);
            mn0 = __vimin3_s16x2(mn0, p[c] * 512u, p[c + 1] * 512u);
            mx1 = __vimax3_s16x2(mx1, p[c + 2], p[c + 3]);
            mn1 = __vimin3_s16x2(mn1, p[c + 2] * 512u, p[c + 3] * 512u);
        }
        return ((mx0 ^ hi_pk) | (mx1 ^ hi_pk) | (mn0 ^ lo_pk) | (mn1 ^ lo_pk)) != 0;
    }
    uint32_t mx = hi_pk, mn = lo_pk;
#pragma unroll
    for (int c = 0; c < NREG; c += 2) {
        mx = __vimax3_s16x2(mx, p[c], p[c + 1]);
        mn = __vimin3_s16x2(mn, p[c] * 512u, p[c + 1] * 512u);
    }
    return ((mx ^ hi_pk) | (mn ^ lo_pk)) != 0;
}

// Cold path of the later schedules: the hot test fired somewhere in this lane's register image.  A fire stalls not just
// this warp but the CTA's pipeline (the MMA issuer needs all sixteen arrivals per stage), so it has to be short: the
// same packed min/max test, one register at a time, finds WHICH registers crossed a bound (no memory access); only their
// four codes each (32 bytes) are re-read from global memory -- bytes the TMA has just pulled through L2 -- and the scan's
// admission rule is applied to their exact distances.  Nothing is decoded from the accumulators, so the x_a = +-64 alias
// needs no special case and the register image does not have to outlive the loop below.
template <int NREG>
__device__ __forceinline__ void hamming_mma_recheck(const uint32_t (&p)[NREG], uint32_t hi_pk, uint32_t lo_pk, uint64_t first_row,
                                                    uint32_t q, const MmaScanArgs &A, const uint4 *s_q, const uint64_t *s_kid) {
    uint64_t fired = 0;   // bit c: register c (accumulator columns 2c, 2c + 1 = codes 4c .. 4c + 3) crossed a bound
#pragma unroll
    for (int c = 0; c < NREG; ++c)
        fired |= (uint64_t)((__vmaxs2(p[c], hi_pk) != hi_pk) | (__vmins2(p[c] * 512u, lo_pk) != lo_pk)) << c;
    const uint4 s = s_q[q];          // {lo, hi, thr, -}
    const uint64_t kid = s_kid[q];
    while (fired) {
        const uint32_t c = __ffsll((long long)fired) - 1;
        fired &= fired - 1;
        const uint64_t r0 = first_row + 4 * c;
        if (r0 >= A.row_end) break;
        uint32_t lo[4] = {0, 0, 0, 0}, hi[4] = {0, 0, 0, 0};
        if (r0 + 3 < A.row_end) {
            const uint4 v0 = *reinterpret_cast<const uint4 *>(A.codes + r0), v1 = *reinterpret_cast<const uint4 *>(A.codes + r0 + 2);
            lo[0] = v0.x; hi[0] = v0.y; lo[1] = v0.z; hi[1] = v0.w; lo[2] = v1.x; hi[2] = v1.y; lo[3] = v1.z; hi[3] = v1.w;
        } else {
#pragma unroll
            for (int h = 0; h < 4; ++h)
                if (r0 + h < A.row_end) { const uint64_t code = A.codes[r0 + h]; lo[h] = (uint32_t)code; hi[h] = (uint32_t)(code >> 32); }
        }
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            const uint64_t r = r0 + h;
            const uint32_t d = __popc(lo[h] ^ s.x) + __popc(hi[h] ^ s.y);
            if (r >= A.row_end || d > s.z) continue;
            const uint64_t id = A.ids ? (d == s.z ? A.ids[r] : 0) : A.id_base + r;
            if (d < s.z || id < kid) {
                cand_append(A.cand, A.count, A.cap, q, ((uint64_t)d << 40) | r);
            }
        }
    }
}

template <int kEpiW>
__global__ void __launch_bounds__(32 * (2 + kEpiW), 1)
hamming_mma_scan2_kernel(const __grid_constant__ MmaScanArgs A) {
    constexpr int kColsW = kMmaRows / (kEpiW / 4);     // accumulator columns per epilogue warp: 64 or 128
    constexpr int kHalfRegs = kColsW / 4;                // packed registers per half: 16 or 32
    constexpr uint32_t kHalfCodes = kColsW;              // codes per half (2 per column, kColsW / 2 columns)
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool spin_ctl = A.wait_flags & 1u, spin_epi = A.wait_flags & 2u;
    auto wait_ctl = [&](uint64_t *bar, uint32_t ph) { if (spin_ctl) mbar_wait(bar, ph); else mbar_wait_sleep(bar, ph); };
    auto wait_epi = [&](uint64_t *bar, uint32_t ph) { if (spin_epi) mbar_wait(bar, ph); else mbar_wait_sleep(bar, ph); };
    const uint32_t q_tiles = (A.nq + kMmaQTile - 1) / kMmaQTile;
    unsigned char *sQ = smem;
    unsigned char *sC = smem + (size_t)q_tiles * kMmaQBytes;
    const uint32_t n_stages = min((uint32_t)kMmaMaxImgStages, 12u - q_tiles);
    uint4 *s_q = reinterpret_cast<uint4 *>(smem + (size_t)(kMmaMaxQueries / kMmaQTile) * kMmaQBytes + kMmaStages * kMmaCBytes);
    uint64_t *s_kid = reinterpret_cast<uint64_t *>(s_q + kMmaMaxQueries);
    uint2 *s_bnd = reinterpret_cast<uint2 *>(s_kid + kMmaMaxQueries);
    uint64_t *cfull = reinterpret_cast<uint64_t *>(s_bnd + kMmaMaxQueries);
    uint64_t *cempty = cfull + kMmaMaxImgStages, *tfull = cempty + kMmaMaxImgStages, *tempty = tfull + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty + 2);
    const uint32_t n_tiles = (uint32_t)((A.row_end - A.row0 + kMmaTileCodes - 1) / kMmaTileCodes);
    const uint32_t my_tiles = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    for (uint32_t q = threadIdx.x; q < q_tiles * kMmaQTile; q += blockDim.x) {
        const QSlot s = q < A.nq ? A.slots[q] : QSlot{0, 0, 0, 0};
        mma_store_query_row(sQ + (q / kMmaQTile) * kMmaQBytes, q % kMmaQTile, s.lo, s.hi, q < A.nq);
        const uint64_t kid = q < A.nq ? A.kth_id[q] : 0;
        s_q[q] = make_uint4(s.lo, s.hi, s.thr, 0);
        s_kid[q] = kid;
        uint32_t hot = s.thr;   // as in hamming_mma_scan_kernel: thr - 1 under implicit ids, "never" for padding queries
        if (A.ids == nullptr && kid < A.id_base + A.row0) hot = s.thr == 0 ? 0xFFFFFFFFu : s.thr - 1;
        if (q >= A.nq) hot = 0xFFFFFFFFu;
        const int32_t tau = 64 - 2 * (int32_t)hot;
        int32_t hi16 = 64 * (tau - 1), lo16 = (-tau * 512) | 0x1FF;
        if (hot == 0xFFFFFFFFu) { hi16 = 0x8000; lo16 = -0x8001; }
        else if (hot >= 64) { hi16 = -0x7FFF; lo16 = 0; }
        const uint32_t hi1 = (uint16_t)(hi16 - 1), lo1 = (uint16_t)(lo16 + 1);
        s_bnd[q] = make_uint2(hi1 << 16 | hi1, lo1 << 16 | lo1);   // hi1 == 0x7FFF only for "never" (kMmaNeverHiPk)
    }
    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < n_stages; ++s) { mbar_init(&cfull[s], 1); mbar_init(&cempty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], kEpiW); }
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc_512(tmem_slot);
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== MMA issuer =====
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kMmaRows >> 3) << 17) | ((uint32_t)(kMmaQTile >> 4) << 24);
        uint32_t acc_it = 0;
        for (uint32_t it = 0; it < my_tiles; ++it) {
            const uint32_t s = it % n_stages, ph = (it / n_stages) & 1;
            wait_ctl(&cfull[s], ph);
            tcgen05_fence_after();
            const uint64_t bdesc = umma_desc_sw64(smem_u32(sC + s * kMmaImgBytes));
            for (uint32_t mt = 0; mt < q_tiles; ++mt, ++acc_it) {
                const uint32_t as = acc_it & 1, aph = (acc_it >> 1) & 1;
                wait_ctl(&tempty[as], aph ^ 1);
                tcgen05_fence_after();
                if (lane == 0) {
                    const uint64_t adesc = umma_desc_sw128(smem_u32(sQ + mt * kMmaQBytes));
                    umma_i8(tmem_base + as * kMmaRows, adesc, bdesc, idesc, 0u);
                    umma_i8(tmem_base + as * kMmaRows, adesc + 2, bdesc + 2, idesc, 1u);
                    umma_commit(&tfull[as]);
                }
                __syncwarp();
            }
            if (lane == 0) umma_commit(&cempty[s]);
            __syncwarp();
        }
    } else if (warp == 1) {
        // ===== producer: one 16 KiB TMA bulk copy per stage image =====
        if (lane == 0) {
            const unsigned char *src = reinterpret_cast<const unsigned char *>(A.ops) + (A.row0 / kMmaTileCodes) * (uint64_t)kMmaImgBytes;
            for (uint32_t it = 0; it < my_tiles; ++it) {
                const uint32_t s = it % n_stages, ph = (it / n_stages) & 1;
                const uint64_t tile = blockIdx.x + (uint64_t)it * gridDim.x;
                wait_ctl(&cempty[s], ph ^ 1);
                mbar_expect_tx(&cfull[s], kMmaImgBytes);
                tma_bulk_g2s(sC + s * kMmaImgBytes, src + tile * kMmaImgBytes, kMmaImgBytes, &cfull[s]);
            }
        }
    } else {
        // ===== epilogue: warp -> TMEM lane quadrant (warp % 4) and kColsW of the 256 columns, read as two halves =====
        const uint32_t quad = warp & 3, part = (uint32_t)(warp - 2) >> 2;
        const uint32_t taddr0 = tmem_base + ((quad * 32u) << 16) + part * kColsW;
        const uint32_t n_items = my_tiles * q_tiles;            // item = (stage tile, query tile); accumulator stage = item & 1
        const uint32_t bnd0 = smem_u32(s_bnd + quad * 32 + lane), bnd_end = bnd0 + q_tiles * (kMmaQTile * 8u);
        uint32_t bnd_at = bnd0;                                  // this thread's bounds for the query tile of the item in hand
        uint32_t par = 0;                                        // mbarrier phase parity of the stage pair in hand
        uint32_t pa[kHalfRegs], pb[kHalfRegs];
        // first row of (item, half h); only the cold path needs it
        auto half_row = [&](uint32_t item, uint32_t h) {
            const uint32_t it = item / q_tiles;
            return A.row0 + ((uint64_t)blockIdx.x + (uint64_t)it * gridDim.x) * kMmaTileCodes + 2 * part * kColsW + h * kHalfCodes;
        };
        auto q_of = [&](uint32_t item) { return (item % q_tiles) * kMmaQTile + quad * 32 + lane; };
        // one item on accumulator stage `stg` (compile-time): pa holds its first half, in flight
        auto do_item = [&](const uint32_t stg, uint32_t item) {
            uint32_t hi_pk, lo_pk;
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(hi_pk), "=r"(lo_pk) : "r"(bnd_at));
            const uint32_t taddr = taddr0 + stg * kMmaRows;
            tmem_ld_wait(pa);                                    // first half landed ...
            tmem_ld_pack16_async(taddr + kColsW / 2, pb);        // ... second half flies while the first is reduced
            if (hamming_mma_hot_test<kHalfRegs>(pa, hi_pk, lo_pk)) hamming_mma_recheck<kHalfRegs>(pa, hi_pk, lo_pk, half_row(item, 0), q_of(item), A, s_q, s_kid);
            tmem_ld_wait(pb);                                    // the whole stage now lives in registers: hand it back
            tcgen05_fence_before();
            if (lane == 0) mbar_arrive(&tempty[stg]);
            if (item + 1 < n_items) {
                wait_epi(&tfull[stg ^ 1], stg ? par ^ 1 : par);
                tcgen05_fence_after();
                tmem_ld_pack16_async(taddr0 + (stg ^ 1) * kMmaRows, pa);   // next item's first half flies while this one's second is reduced
            }
            if (hamming_mma_hot_test<kHalfRegs>(pb, hi_pk, lo_pk)) hamming_mma_recheck<kHalfRegs>(pb, hi_pk, lo_pk, half_row(item, 1), q_of(item), A, s_q, s_kid);
            bnd_at += kMmaQTile * 8u;
            if (bnd_at == bnd_end) bnd_at = bnd0;
        };
        if (n_items) {
            wait_epi(&tfull[0], 0);
            tcgen05_fence_after();
            tmem_ld_pack16_async(taddr0, pa);
            for (uint32_t item = 0; item < n_items; item += 2) {
                do_item(0, item);
                if (item + 1 >= n_items) break;
                do_item(1, item + 1);
                par ^= 1;
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) {
        tcgen05_fence_after();
        tmem_dealloc_512(tmem_base);
    }
}

// ---- third schedule: one 64-column load per item as in the first generation, but the NEXT item's load is issued before
// the current item is reduced (two 32-register images per epilogue thread, 18 warps so that ptxas may use 96 registers).
// The stage is handed back at the top of the step, as early as in the first generation, so the MMA issuer has a whole
// epilogue period to produce the next accumulator; what disappears is the exposed tcgen05.ld latency that all sixteen
// epilogue warps used to sit out together once per item.
template <int kEpiW>
__global__ void __launch_bounds__(32 * (2 + kEpiW), 1)
hamming_mma_scan3_kernel(const __grid_constant__ MmaScanArgs A) {
    constexpr int kColsW = kMmaRows / (kEpiW / 4);   // accumulator columns per epilogue warp: 64 (16 warps) or 128 (8 warps)
    constexpr int kRegs = kColsW / 2;                  // packed registers per item: 32 or 64
    // (a 20-warp form that moved registers from the control warps to the epilogue warps with setmaxnreg hung on the device and was dropped)
    constexpr int kFirstEpiWarp = 2;
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool spin_ctl = A.wait_flags & 1u, spin_epi = A.wait_flags & 2u;
    auto wait_ctl = [&](uint64_t *bar, uint32_t ph) { if (spin_ctl) mbar_wait(bar, ph); else mbar_wait_sleep(bar, ph); };
    auto wait_epi = [&](uint64_t *bar, uint32_t ph) { if (spin_epi) mbar_wait(bar, ph); else mbar_wait_sleep(bar, ph); };
    const uint32_t q_tiles = (A.nq + kMmaQTile - 1) / kMmaQTile;
    unsigned char *sQ = smem;
    unsigned char *sC = smem + (size_t)q_tiles * kMmaQBytes;
    const uint32_t n_stages = min((uint32_t)kMmaMaxImgStages, 12u - q_tiles);
    uint4 *s_q = reinterpret_cast<uint4 *>(smem + (size_t)(kMmaMaxQueries / kMmaQTile) * kMmaQBytes + kMmaStages * kMmaCBytes);
    uint64_t *s_kid = reinterpret_cast<uint64_t *>(s_q + kMmaMaxQueries);
    uint2 *s_bnd = reinterpret_cast<uint2 *>(s_kid + kMmaMaxQueries);
    uint64_t *cfull = reinterpret_cast<uint64_t *>(s_bnd + kMmaMaxQueries);
    uint64_t *cempty = cfull + kMmaMaxImgStages, *tfull = cempty + kMmaMaxImgStages, *tempty = tfull + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty + 2);
    const uint32_t n_tiles = (uint32_t)((A.row_end - A.row0 + kMmaTileCodes - 1) / kMmaTileCodes);
    const uint32_t my_tiles = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    for (uint32_t q = threadIdx.x; q < q_tiles * kMmaQTile; q += blockDim.x) {
        const QSlot s = q < A.nq ? A.slots[q] : QSlot{0, 0, 0, 0};
        mma_store_query_row(sQ + (q / kMmaQTile) * kMmaQBytes, q % kMmaQTile, s.lo, s.hi, q < A.nq);
        const uint64_t kid = q < A.nq ? A.kth_id[q] : 0;
        s_q[q] = make_uint4(s.lo, s.hi, s.thr, 0);
        s_kid[q] = kid;
        uint32_t hot = s.thr;   // as in hamming_mma_scan_kernel
        if (A.ids == nullptr && kid < A.id_base + A.row0) hot = s.thr == 0 ? 0xFFFFFFFFu : s.thr - 1;
        if (q >= A.nq) hot = 0xFFFFFFFFu;
        const int32_t tau = 64 - 2 * (int32_t)hot;
        int32_t hi16 = 64 * (tau - 1), lo16 = (-tau * 512) | 0x1FF;
        if (hot == 0xFFFFFFFFu) { hi16 = 0x8000; lo16 = -0x8001; }
        else if (hot >= 64) { hi16 = -0x7FFF; lo16 = 0; }
        const uint32_t hi1 = (uint16_t)(hi16 - 1), lo1 = (uint16_t)(lo16 + 1);
        s_bnd[q] = make_uint2(hi1 << 16 | hi1, lo1 << 16 | lo1);
    }
    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < n_stages; ++s) { mbar_init(&cfull[s], 1); mbar_init(&cempty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], kEpiW); }
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc_512(tmem_slot);
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < kFirstEpiWarp) {
    if (warp == 0) {
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kMmaRows >> 3) << 17) | ((uint32_t)(kMmaQTile >> 4) << 24);
        uint32_t acc_it = 0;
        for (uint32_t it = 0; it < my_tiles; ++it) {
            const uint32_t s = it % n_stages, ph = (it / n_stages) & 1;
            wait_ctl(&cfull[s], ph);
            tcgen05_fence_after();
            const uint64_t bdesc = umma_desc_sw64(smem_u32(sC + s * kMmaImgBytes));
            for (uint32_t mt = 0; mt < q_tiles; ++mt, ++acc_it) {
                const uint32_t as = acc_it & 1, aph = (acc_it >> 1) & 1;
                wait_ctl(&tempty[as], aph ^ 1);
                tcgen05_fence_after();
                if (lane == 0) {
                    const uint64_t adesc = umma_desc_sw128(smem_u32(sQ + mt * kMmaQBytes));
                    umma_i8(tmem_base + as * kMmaRows, adesc, bdesc, idesc, 0u);
                    umma_i8(tmem_base + as * kMmaRows, adesc + 2, bdesc + 2, idesc, 1u);
                    umma_commit(&tfull[as]);
                }
                __syncwarp();
            }
            if (lane == 0) umma_commit(&cempty[s]);
            __syncwarp();
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const unsigned char *src = reinterpret_cast<const unsigned char *>(A.ops) + (A.row0 / kMmaTileCodes) * (uint64_t)kMmaImgBytes;
            for (uint32_t it = 0; it < my_tiles; ++it) {
                const uint32_t s = it % n_stages, ph = (it / n_stages) & 1;
                const uint64_t tile = blockIdx.x + (uint64_t)it * gridDim.x;
                wait_ctl(&cempty[s], ph ^ 1);
                mbar_expect_tx(&cfull[s], kMmaImgBytes);
                tma_bulk_g2s(sC + s * kMmaImgBytes, src + tile * kMmaImgBytes, kMmaImgBytes, &cfull[s]);
            }
        }
    }
    } else {
        const uint32_t quad = warp & 3, part = (uint32_t)(warp - kFirstEpiWarp) >> 2;
        const uint32_t taddr0 = tmem_base + ((quad * 32u) << 16) + part * kColsW;
        const uint32_t n_items = my_tiles * q_tiles;
        const uint32_t bnd0 = smem_u32(s_bnd + quad * 32 + lane), bnd_end = bnd0 + q_tiles * (kMmaQTile * 8u);
        uint32_t bnd_at = bnd0, par = 0;
        uint32_t pa[kRegs], pb[kRegs];
        auto item_row = [&](uint32_t item) {
            const uint32_t it = item / q_tiles;
            return A.row0 + ((uint64_t)blockIdx.x + (uint64_t)it * gridDim.x) * kMmaTileCodes + 2 * part * kColsW;
        };
        auto step = [&](uint32_t (&cur)[kRegs], uint32_t (&nxt)[kRegs], const uint32_t stg, uint32_t item) {
            uint32_t hi_pk, lo_pk;
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(hi_pk), "=r"(lo_pk) : "r"(bnd_at));
            tmem_ld_wait(cur);                                   // the only outstanding load of this thread
            tcgen05_fence_before();
            if (lane == 0) mbar_arrive(&tempty[stg]);            // accumulators are in registers: hand the stage back at once
            if (item + 1 < n_items) {
                wait_epi(&tfull[stg ^ 1], stg ? par ^ 1 : par);
                tcgen05_fence_after();
                tmem_ld_pack16_async(taddr0 + (stg ^ 1) * kMmaRows, nxt);   // lands while `cur` is reduced
            }
            if (hamming_mma_hot_test<kRegs>(cur, hi_pk, lo_pk))
                hamming_mma_recheck<kRegs>(cur, hi_pk, lo_pk, item_row(item), (item % q_tiles) * kMmaQTile + quad * 32 + lane, A, s_q, s_kid);
            bnd_at += kMmaQTile * 8u;
            if (bnd_at == bnd_end) bnd_at = bnd0;
        };
        if (n_items) {
            wait_epi(&tfull[0], 0);
            tcgen05_fence_after();
            tmem_ld_pack16_async(taddr0, pa);
            for (uint32_t item = 0; item < n_items; item += 2) {
                step(pa, pb, 0, item);
                if (item + 1 >= n_items) break;
                step(pb, pa, 1, item + 1);
                par ^= 1;
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) {
        tcgen05_fence_after();
        tmem_dealloc_512(tmem_base);
    }
}

// ---- fourth schedule: FOUR accumulator stages of 128 columns (N = 128 MMAs) ---------------------------------------------
// Each epilogue warp reads its slice of a stage with ONE tcgen05.ld (kEpiW = 8: 64 columns / 32 registers; 16: 32 columns /
// 16 registers), hands the stage back the moment that load has landed -- as early as the first generation -- and has the
// NEXT stage's load in flight while it reduces the current one.  With four stages the MMA issuer runs up to three items
// ahead, so neither side waits on a hand-over in steady state.
template <int kEpiW>
__global__ void __launch_bounds__(32 * (2 + kEpiW), 1)
hamming_mma_scan4_kernel(const __grid_constant__ MmaScanArgs A) {
    constexpr int kStageCols = 128;                       // accumulator stage = half of a 256-row operand stage
    constexpr int kColsW = kStageCols / (kEpiW / 4);      // columns per epilogue warp: 64 or 32
    constexpr int kRegs = kColsW / 2;                     // packed registers per item: 32 or 16
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool spin_ctl = A.wait_flags & 1u, spin_epi = A.wait_flags & 2u;
    auto wait_ctl = [&](uint64_t *bar, uint32_t ph) { if (spin_ctl) mbar_wait(bar, ph); else mbar_wait_sleep(bar, ph); };
    auto wait_epi = [&](uint64_t *bar, uint32_t ph) { if (spin_epi) mbar_wait(bar, ph); else mbar_wait_sleep(bar, ph); };
    const uint32_t q_tiles = (A.nq + kMmaQTile - 1) / kMmaQTile;
    unsigned char *sQ = smem;
    unsigned char *sC = smem + (size_t)q_tiles * kMmaQBytes;
    const uint32_t n_stages = min((uint32_t)kMmaMaxImgStages, 12u - q_tiles);
    uint4 *s_q = reinterpret_cast<uint4 *>(smem + (size_t)(kMmaMaxQueries / kMmaQTile) * kMmaQBytes + kMmaStages * kMmaCBytes);
    uint64_t *s_kid = reinterpret_cast<uint64_t *>(s_q + kMmaMaxQueries);
    uint2 *s_bnd = reinterpret_cast<uint2 *>(s_kid + kMmaMaxQueries);
    uint64_t *cfull = reinterpret_cast<uint64_t *>(s_bnd + kMmaMaxQueries);
    uint64_t *cempty = cfull + kMmaMaxImgStages, *tfull = cempty + kMmaMaxImgStages, *tempty = tfull + 4;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty + 4);
    const uint32_t n_tiles = (uint32_t)((A.row_end - A.row0 + kMmaTileCodes - 1) / kMmaTileCodes);
    const uint32_t my_tiles = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    for (uint32_t q = threadIdx.x; q < q_tiles * kMmaQTile; q += blockDim.x) {
        const QSlot s = q < A.nq ? A.slots[q] : QSlot{0, 0, 0, 0};
        mma_store_query_row(sQ + (q / kMmaQTile) * kMmaQBytes, q % kMmaQTile, s.lo, s.hi, q < A.nq);
        const uint64_t kid = q < A.nq ? A.kth_id[q] : 0;
        s_q[q] = make_uint4(s.lo, s.hi, s.thr, 0);
        s_kid[q] = kid;
        uint32_t hot = s.thr;   // as in hamming_mma_scan_kernel
        if (A.ids == nullptr && kid < A.id_base + A.row0) hot = s.thr == 0 ? 0xFFFFFFFFu : s.thr - 1;
        if (q >= A.nq) hot = 0xFFFFFFFFu;
        const int32_t tau = 64 - 2 * (int32_t)hot;
        int32_t hi16 = 64 * (tau - 1), lo16 = (-tau * 512) | 0x1FF;
        if (hot == 0xFFFFFFFFu) { hi16 = 0x8000; lo16 = -0x8001; }
        else if (hot >= 64) { hi16 = -0x7FFF; lo16 = 0; }
        const uint32_t hi1 = (uint16_t)(hi16 - 1), lo1 = (uint16_t)(lo16 + 1);
        s_bnd[q] = make_uint2(hi1 << 16 | hi1, lo1 << 16 | lo1);
    }
    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < n_stages; ++s) { mbar_init(&cfull[s], 1); mbar_init(&cempty[s], 1); }
        for (int s = 0; s < 4; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], kEpiW); }
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc_512(tmem_slot);
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== MMA issuer: item = (stage tile, query tile, half): D[128 queries x 128 rows], two K = 32 steps =====
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kStageCols >> 3) << 17) | ((uint32_t)(kMmaQTile >> 4) << 24);
        uint32_t acc_it = 0;
        for (uint32_t it = 0; it < my_tiles; ++it) {
            const uint32_t s = it % n_stages, ph = (it / n_stages) & 1;
            wait_ctl(&cfull[s], ph);
            tcgen05_fence_after();
            const uint64_t bdesc0 = umma_desc_sw64(smem_u32(sC + s * kMmaImgBytes));
            for (uint32_t mt = 0; mt < q_tiles; ++mt)
#pragma unroll
                for (uint32_t h = 0; h < 2; ++h, ++acc_it) {
                    const uint32_t as = acc_it & 3, aph = (acc_it >> 2) & 1;
                    wait_ctl(&tempty[as], aph ^ 1);
                    tcgen05_fence_after();
                    if (lane == 0) {
                        const uint64_t adesc = umma_desc_sw128(smem_u32(sQ + mt * kMmaQBytes));
                        const uint64_t bdesc = bdesc0 + h * ((kStageCols * 64) >> 4);   // operand rows 128 h .. of the stage image
                        umma_i8(tmem_base + as * kStageCols, adesc, bdesc, idesc, 0u);
                        umma_i8(tmem_base + as * kStageCols, adesc + 2, bdesc + 2, idesc, 1u);
                        umma_commit(&tfull[as]);
                    }
                    __syncwarp();
                }
            if (lane == 0) umma_commit(&cempty[s]);
            __syncwarp();
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const unsigned char *src = reinterpret_cast<const unsigned char *>(A.ops) + (A.row0 / kMmaTileCodes) * (uint64_t)kMmaImgBytes;
            for (uint32_t it = 0; it < my_tiles; ++it) {
                const uint32_t s = it % n_stages, ph = (it / n_stages) & 1;
                const uint64_t tile = blockIdx.x + (uint64_t)it * gridDim.x;
                wait_ctl(&cempty[s], ph ^ 1);
                mbar_expect_tx(&cfull[s], kMmaImgBytes);
                tma_bulk_g2s(sC + s * kMmaImgBytes, src + tile * kMmaImgBytes, kMmaImgBytes, &cfull[s]);
            }
        }
    } else {
        // ===== epilogue =====
        const uint32_t quad = warp & 3, part = (uint32_t)(warp - 2) >> 2;
        const uint32_t taddr0 = tmem_base + ((quad * 32u) << 16) + part * kColsW;
        const uint32_t n_items = my_tiles * q_tiles * 2;
        const uint32_t bnd0 = smem_u32(s_bnd + quad * 32 + lane), bnd_end = bnd0 + q_tiles * (kMmaQTile * 8u);
        uint32_t bnd_at = bnd0;
        uint32_t pa[kRegs], pb[kRegs];
        // first code of (item, this warp's columns): tile, half of the operand stage, column slice (two codes per column)
        auto item_row = [&](uint32_t item) {
            const uint32_t it = item / (2 * q_tiles), h = item & 1;
            return A.row0 + ((uint64_t)blockIdx.x + (uint64_t)it * gridDim.x) * kMmaTileCodes + 2 * (h * kStageCols + part * kColsW);
        };
        auto step = [&](uint32_t (&cur)[kRegs], uint32_t (&nxt)[kRegs], uint32_t item) {
            uint32_t hi_pk, lo_pk;
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(hi_pk), "=r"(lo_pk) : "r"(bnd_at));
            tmem_ld_wait(cur);                                   // the only outstanding load of this thread
            tcgen05_fence_before();
            if (lane == 0) mbar_arrive(&tempty[item & 3]);       // accumulators are in registers: hand the stage back at once
            if (item + 1 < n_items) {
                wait_epi(&tfull[(item + 1) & 3], ((item + 1) >> 2) & 1);
                tcgen05_fence_after();
                tmem_ld_pack16_async(taddr0 + ((item + 1) & 3) * kStageCols, nxt);   // lands while `cur` is reduced
            }
            if (hamming_mma_hot_test<kRegs>(cur, hi_pk, lo_pk))
                hamming_mma_recheck<kRegs>(cur, hi_pk, lo_pk, item_row(item), ((item >> 1) % q_tiles) * kMmaQTile + quad * 32 + lane, A, s_q, s_kid);
            if (item & 1) { bnd_at += kMmaQTile * 8u; if (bnd_at == bnd_end) bnd_at = bnd0; }   // next query tile after both halves
        };
        if (n_items) {
            wait_epi(&tfull[0], 0);
            tcgen05_fence_after();
            tmem_ld_pack16_async(taddr0, pa);
            for (uint32_t item = 0; item < n_items; item += 2) {   // n_items is even
                step(pa, pb, item);
                step(pb, pa, item + 1);
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) {
        tcgen05_fence_after();
        tmem_dealloc_512(tmem_base);
    }
}

