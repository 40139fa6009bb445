// Bulk path of the SSE step: k_prime -> k_relay -> k_commit.   (included from sse_kernels.cuh)
//
// Regular streams -- committed, carries empty and equal at the start of the bulk region, every
// chunk valid UTF-8 and non-empty, no run of three or more LFs, no event the tap turns into an
// extra row ("error") or that has an unmodelled shape -- are exactly the streams for which the
// reference's chunk-by-chunk split (request_handler.py:111-115, chat_logging.py:108-112) equals a
// split of the CONCATENATED text on every LF LF pair.  For those, events are independent: the
// chunk in which an event completes owns it, parses it, and posts its findings with atomics.
// Everything else is flagged irregular and redone by k_commit with the exact sequential machine,
// so the result never depends on which path ran (tests run both and compare with the oracle).


#define LGW_RELAY_THREADS 256
#define LGW_TILE_VECS (LGW_TILE_BYTES / 16)

// 16 KB tile in shared memory, swizzled so that 32 lanes reading the same byte position of 32
// consecutive 64-byte rows (the 64-byte-event pattern) hit 32 different banks, while a 16-byte
// vector store stays one STS.128 (the four words are permuted inside their own vector).
__device__ __forceinline__ uint32_t swz(uint32_t off) {
    const uint32_t r = off >> 6;
    return (off & ~63u) | ((((off >> 4) & 3u) ^ ((r >> 1) & 3u)) << 4) | ((((off >> 2) & 3u) ^ ((r >> 3) & 3u)) << 2) | (off & 3u);
}

// per-byte equality of a 32-bit word with a repeated byte -> 4-bit mask (bit k = byte k equal)
__device__ __forceinline__ uint32_t eq4(uint32_t x, uint32_t pat4) {
    const uint32_t y = x ^ pat4;
    const uint32_t t = ~(((y & 0x7f7f7f7fu) + 0x7f7f7f7fu) | y | 0x7f7f7f7fu);     // 0x80 where the byte is zero
    return ((t >> 7) * 0x01020408u) >> 24;
}
__device__ __forceinline__ uint32_t lf_mask16(const uint4& x) {
    return eq4(x.x, 0x0a0a0a0au) | (eq4(x.y, 0x0a0a0a0au) << 4) | (eq4(x.z, 0x0a0a0a0au) << 8) | (eq4(x.w, 0x0a0a0a0au) << 12);
}

struct TileReader {
    const uint8_t* smem;           // swizzled tile
    const uint32_t* vinfo;         // per 16-byte vector: LF mask | (has byte >= 0x80) << 16
    const uint8_t* __restrict__ g; // whole packed buffer
    uint32_t t0, n_bytes;
    __device__ __forceinline__ uint32_t at(uint32_t pos) const {
        const uint32_t d = pos - t0;
        return d < LGW_TILE_BYTES ? (uint32_t)smem[swz(d)] : (uint32_t)__ldg(g + pos);
    }
    // aligned 32-bit word containing byte `pos` (little endian); bytes past n_bytes read as 0
    __device__ __forceinline__ uint32_t word(uint32_t pos) const {
        const uint32_t p4 = pos & ~3u, d = p4 - t0;
        if (d < LGW_TILE_BYTES) return *reinterpret_cast<const uint32_t*>(smem + swz(d));
        if (p4 + 4 <= n_bytes) return __ldg(reinterpret_cast<const uint32_t*>(g + p4));
        uint32_t w = 0;
        for (uint32_t k = 0; k < 4 && p4 + k < n_bytes; ++k) w |= (uint32_t)__ldg(g + p4 + k) << (8 * k);
        return w;
    }
    // LF mask / high flag of the 16-byte vector number V (= byte offset >> 4)
    __device__ __forceinline__ uint32_t vec_info(uint32_t V) const {
        const uint32_t d = V - (t0 >> 4);
        if (d < LGW_TILE_VECS) return vinfo[d];
        const uint32_t p = V << 4;
        if (p >= n_bytes) return 0;
        uint32_t m = 0, hi = 0;
        if (p + 16 <= n_bytes) {
            const uint4 x = __ldg(reinterpret_cast<const uint4*>(g + p));
            m = lf_mask16(x); hi = ((x.x | x.y | x.z | x.w) & 0x80808080u) ? 1u : 0u;
        } else {
            for (uint32_t k = 0; p + k < n_bytes; ++k) { const uint32_t c = __ldg(g + p + k); if (c == '\n') m |= 1u << k; if (c >= 0x80) hi = 1; }
        }
        return m | (hi << 16);
    }
    __device__ __forceinline__ bool lf_at(uint32_t pos) const { return (vec_info(pos >> 4) >> (pos & 15)) & 1u; }
};

// rare path: chunk-level UTF-8 validation (request_handler.py:111 decodes each chunk on its own)
__device__ __noinline__ bool chunk_utf8_ok(const TileReader* rd, uint32_t o, uint32_t e) {
    uint32_t p = o;
    while (p < e) {
        const uint32_t ch = rd->at(p);
        if (ch < 0x80) { ++p; continue; }
        uint32_t need, l = 0x80, h = 0xBF;
        if (ch >= 0xC2 && ch <= 0xDF) need = 1;
        else if (ch == 0xE0) { need = 2; l = 0xA0; }
        else if (ch >= 0xE1 && ch <= 0xEC) need = 2;
        else if (ch == 0xED) { need = 2; h = 0x9F; }
        else if (ch >= 0xEE && ch <= 0xEF) need = 2;
        else if (ch == 0xF0) { need = 3; l = 0x90; }
        else if (ch >= 0xF1 && ch <= 0xF3) need = 3;
        else if (ch == 0xF4) { need = 3; h = 0x8F; }
        else return false;
        if (p + need >= e) return false;
        uint32_t b = rd->at(p + 1);
        if (b < l || b > h) return false;
        for (uint32_t k = 2; k <= need; ++k) { b = rd->at(p + k); if (b < 0x80 || b > 0xBF) return false; }
        p += need + 1;
    }
    return true;
}

// rare path: where does the event that is open at byte o begin?  (o is not right after a separator)
// returns false when the stream must go to the sequential path
__device__ __noinline__ bool find_open_event_start(const TileReader* rd, uint32_t o, uint32_t relay_begin, uint32_t seg_end, uint32_t carry_cap, uint32_t* out_b) {
    uint32_t k = o;
    bool found = false;
    const uint32_t limit = (o - relay_begin > carry_cap + 2) ? o - carry_cap - 2 : relay_begin;
    while (k >= limit + 2) {
        if (rd->at(k - 1) == '\n' && rd->at(k - 2) == '\n') { found = true; break; }
        --k;
    }
    uint32_t b;
    if (found) {
        b = k;
        if (k >= relay_begin + 3 && rd->at(k - 3) == '\n') return false;          // LF run >= 3
        if (b < seg_end && rd->at(b) == '\n') return false;                        // LF run >= 3
    } else {
        b = relay_begin;
        if (limit != relay_begin) return false;                                    // open event longer than the carry capacity
    }
    if (o - b > carry_cap) return false;
    *out_b = b;
    return true;
}

// ---- k_prime ---------------------------------------------------------------------------------------
// thread i: (a) tile table entry i = first chunk that starts at or after byte i*TILE;
//           (b) segment i: finish priming with the exact machine, then plan the bulk region.
__global__ void __launch_bounds__(64) k_prime(StepArgs a, uint32_t n_tiles) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n_tiles) {
        const uint32_t target = i * LGW_TILE_BYTES;
        uint32_t lo = 0, hi = a.n_chunks;                 // first c in [0, n_chunks] with chunk_off[c] >= target
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (__ldg(a.chunk_off + mid) < target) lo = mid + 1; else hi = mid; }
        a.s.tile_chunk[i] = lo;
    }
    if (i >= a.n_segs) return;
    const uint32_t seg = i, c0 = a.seg_chunk[seg], c1 = a.seg_chunk[seg + 1];
    const uint32_t slot = a.seg_slot[seg];
    StreamHdr st = a.t.state[slot].h;                      // local copy of the 64 hot bytes
    SegPlan p;
    p.seg_end = a.chunk_off[c1]; p.relay_begin = p.seg_end; p.irregular = 0; p.last_usage_pos = 0; p.a_usage = 0;
    p.n_events_a = p.n_events_b = p.n_usage_b = 0; p._pad[0] = p._pad[1] = 0;
    uint32_t emit_begin = (st.phase == PH_COMMITTED) ? c0 : c1;
    uint32_t resume = c0;
    if (st.phase == PH_PRIMING) {
        const StepIO io = make_io(a, slot, &st);
        resume = run_chunks(io, a.data, a.chunk_off, c0, c1, emit_begin, true);
        a.t.state[slot].h = st;
    }
    if (st.phase == PH_COMMITTED && resume < c1) {
        if ((st.flags & SF_SYNCED) && st.carry_a_len == 0) p.relay_begin = a.chunk_off[resume];
        else p.irregular = 1;
    }
    p.resume_chunk = resume; p.emit_chunk_begin = emit_begin;
    a.s.plan[seg] = p;
}

// ---- k_relay ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(LGW_RELAY_THREADS, 4) k_relay(StepArgs a) {
    __shared__ __align__(16) uint8_t tile[LGW_TILE_BYTES];
    __shared__ uint32_t vinfo[LGW_TILE_VECS];
    __shared__ __align__(4) uint8_t s_cls[256];
    __shared__ __align__(4) uint8_t s_trans[LGW_LEAN_ROWS * 32];
    __shared__ uint32_t s_seg_lo, s_seg_hi;

    const uint32_t t0 = blockIdx.x * LGW_TILE_BYTES;
    const uint32_t tid = threadIdx.x;
    const uint32_t n_bytes = a.n_bytes;
    const uint32_t c_lo = a.s.tile_chunk[blockIdx.x], c_hi = a.s.tile_chunk[blockIdx.x + 1];

    // (0) tables + segment range of this tile's chunks
    if (tid < 64) reinterpret_cast<uint32_t*>(s_cls)[tid] = reinterpret_cast<const uint32_t*>(g_lean_tables_dev.cls)[tid];
    else if (tid < 64 + LGW_LEAN_ROWS * 8) reinterpret_cast<uint32_t*>(s_trans)[tid - 64] = reinterpret_cast<const uint32_t*>(g_lean_tables_dev.trans)[tid - 64];
    if (tid >= 224 && tid < 226 && c_hi > c_lo) {
        const uint32_t c = tid == 224 ? c_lo : c_hi - 1;
        uint32_t lo = 0, hi = a.n_segs;
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (__ldg(a.seg_chunk + mid + 1) <= c) lo = mid + 1; else hi = mid; }
        if (tid == 224) s_seg_lo = lo; else s_seg_hi = lo;
    }

    // (1) re-emit: position-preserving 16-byte copy of the tile, staged into shared memory on the
    //     way; LF bit mask and high-byte flag of every vector for the event discovery below
    {
        const uint4* __restrict__ src = reinterpret_cast<const uint4*>(a.data + t0);
        uint4* __restrict__ dst = reinterpret_cast<uint4*>(a.out + t0);
        uint4 x[LGW_TILE_VECS / LGW_RELAY_THREADS];
#pragma unroll
        for (uint32_t k = 0; k < LGW_TILE_VECS / LGW_RELAY_THREADS; ++k) {
            const uint32_t v = k * LGW_RELAY_THREADS + tid;
            x[k] = (t0 + v * 16 + 16 <= n_bytes) ? __ldg(src + v) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (uint32_t k = 0; k < LGW_TILE_VECS / LGW_RELAY_THREADS; ++k) {
            const uint32_t v = k * LGW_RELAY_THREADS + tid;
            const uint32_t pos = t0 + v * 16;
            if (pos + 16 <= n_bytes) {
                dst[v] = x[k];
            } else if (pos < n_bytes) {                        // ragged end of the buffer
                uint32_t w[4] = {0, 0, 0, 0};
                for (uint32_t b = pos; b < n_bytes; ++b) { const uint32_t c = a.data[b]; a.out[b] = (uint8_t)c; w[(b - pos) >> 2] |= c << (8 * ((b - pos) & 3)); }
                x[k] = make_uint4(w[0], w[1], w[2], w[3]);
            }
            const uint32_t r = v >> 2;                          // 64-byte row
            const uint32_t kx = (r >> 3) & 3u;                  // word permutation inside the vector
            uint4 y;
            y.x = kx == 0 ? x[k].x : kx == 1 ? x[k].y : kx == 2 ? x[k].z : x[k].w;
            y.y = kx == 0 ? x[k].y : kx == 1 ? x[k].x : kx == 2 ? x[k].w : x[k].z;
            y.z = kx == 0 ? x[k].z : kx == 1 ? x[k].w : kx == 2 ? x[k].x : x[k].y;
            y.w = kx == 0 ? x[k].w : kx == 1 ? x[k].z : kx == 2 ? x[k].y : x[k].x;
            const uint32_t slot16 = (r << 6) | (((v & 3u) ^ ((r >> 1) & 3u)) << 4);
            *reinterpret_cast<uint4*>(tile + slot16) = y;
            vinfo[v] = lf_mask16(x[k]) | ((((x[k].x | x[k].y | x[k].z | x[k].w) & 0x80808080u) ? 1u : 0u) << 16);
        }
    }
    __syncthreads();

    // (2) events of the chunks that START in this tile
    const TileReader rd{tile, vinfo, a.data, t0, n_bytes};
    const uint32_t seg_lo = s_seg_lo, seg_hi = s_seg_hi;
    uint32_t acc_seg = 0xFFFFFFFFu, ev_a = 0, ev_b = 0;       // per-thread counters of the current segment

    for (uint32_t c = c_lo + tid; c < c_hi; c += LGW_RELAY_THREADS) {
        uint32_t seg = seg_lo;
        if (seg_lo != seg_hi) {                                // segment of chunk c: last seg with seg_chunk[seg] <= c
            uint32_t lo = seg_lo, hi = seg_hi;
            while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (__ldg(a.seg_chunk + mid + 1) <= c) lo = mid + 1; else hi = mid; }
            seg = lo;
        }
        SegPlan* pl = a.s.plan + seg;
        const uint32_t relay_begin = pl->relay_begin, seg_end = pl->seg_end;
        const uint32_t o = __ldg(a.chunk_off + c), e = __ldg(a.chunk_off + c + 1);
        if (o < relay_begin || pl->irregular) continue;
        if (e == o) { pl->irregular = 1; continue; }
        if (seg != acc_seg) {
            if (acc_seg != 0xFFFFFFFFu) { if (ev_a) atomicAdd(&a.s.plan[acc_seg].n_events_a, ev_a); if (ev_b) atomicAdd(&a.s.plan[acc_seg].n_events_b, ev_b); }
            acc_seg = seg; ev_a = ev_b = 0;
        }

        // chunk-level UTF-8: only when some vector of the chunk has a byte >= 0x80
        {
            uint32_t hi_any = 0;
            for (uint32_t V = o >> 4; V <= (e - 1) >> 4; ++V) hi_any |= rd.vec_info(V) >> 16;
            if (hi_any && !chunk_utf8_ok(&rd, o, e)) { pl->irregular = 1; continue; }
        }

        // where does the event that is open at the start of this chunk begin?
        uint32_t b = o;
        if (o != relay_begin) {
            const bool sep_before = o >= relay_begin + 2 && rd.lf_at(o - 1) && rd.lf_at(o - 2);
            if (sep_before) {
                if ((o >= relay_begin + 3 && rd.lf_at(o - 3)) || rd.lf_at(o)) { pl->irregular = 1; continue; }   // LF run >= 3
            } else if (!find_open_event_start(&rd, o, relay_begin, seg_end, a.t.carry_cap, &b)) { pl->irregular = 1; continue; }
        }

        // separators whose second LF lies in [o, e): first LF i in [max(b, o-1), e-2]
        bool irregular = false;
        uint32_t us_b = 0, a_usage = 0, last_usage = 0;
        uint32_t ps = b;
        const uint32_t i_min = (o > b) ? o - 1 : b;
        if (e >= 2 && i_min + 2 <= e) {
            const uint32_t i_max = e - 2;
            for (uint32_t V = i_min >> 4; V <= (i_max >> 4) && !irregular; ++V) {
                const uint32_t m = (rd.vec_info(V) & 0xFFFFu) | ((rd.vec_info(V + 1) & 3u) << 16);
                uint32_t pairs = m & (m >> 1) & 0xFFFFu;
                const uint32_t base = V << 4;
                if (base < i_min) pairs &= ~((1u << (i_min - base)) - 1u);
                if (base + 15 > i_max) pairs &= (2u << (i_max - base)) - 1u;
                const uint32_t triples = pairs & (m >> 2);
                while (pairs) {
                    const uint32_t k = __ffs(pairs) - 1; pairs &= pairs - 1;
                    const uint32_t i = base + k;
                    if (((triples >> k) & 1u) && i + 2 < seg_end) { irregular = true; break; }
                    if (i < ps) continue;                       // second half of an overlapping pair cannot happen without a triple
                    // ---- one complete event [ps, i) ----
                    const uint32_t len = i - ps;
                    uint32_t cls = PC_NONE;
                    if (len >= 1) {
                        const uint32_t w0 = rd.word(ps) >> (8 * (ps & 3));
                        if ((w0 & 0xff) == '{') cls = PC_BRACE;
                        else if ((w0 & 0xff) == 'd' && len >= 7) {
                            if ((ps & 3) == 0) cls = (w0 == 0x61746164u && (rd.word(ps + 4) & 0xFFFFFFu) == 0x7b203au) ? PC_DATA : PC_NONE;
                            else cls = (rd.at(ps + 1) == 'a' && rd.at(ps + 2) == 't' && rd.at(ps + 3) == 'a' && rd.at(ps + 4) == ':' && rd.at(ps + 5) == ' ' && rd.at(ps + 6) == '{') ? PC_DATA : PC_NONE;
                        }
                    }
                    if (cls != PC_NONE) {
                        LeanMachine lm;
                        lm.reset(cls == PC_DATA);
                        uint32_t pos = ps + (cls == PC_DATA ? 6u : 0u);
                        while (pos < i) {                        // word-wise byte loop
                            uint32_t w = rd.word(pos) >> (8 * (pos & 3));
                            uint32_t nb = 4 - (pos & 3);
                            if (nb > i - pos) nb = i - pos;
#pragma unroll 1
                            for (uint32_t q = 0; q < nb; ++q, ++pos, w >>= 8) lm.step(w & 0xffu, pos, rd, s_cls, s_trans);
                        }
                        uint32_t f = lm.finish();
                        if (cls == PC_DATA) {                    // handler loop, request_handler.py:122-134
                            ++ev_a;
                            if ((f & PF_VALID_A) && !(f & TK_CODE) && (f & TK_USAGE)) a_usage = 1;
                        }
                        if (f & PF_VALID_B) {                    // tap loop, chat_logging.py:123-141
                            ++ev_b;
                            // events with "error" (extra DB row) go to the sequential path; a "usage" event is
                            // a candidate that k_commit validates with the full machine (choices walk)
                            if (f & TK_ERROR) { irregular = true; break; }
                            if (f & TK_USAGE) { ++us_b; last_usage = ps + 1; }
                        }
                    }
                    ps = i + 2;
                }
            }
        }
        if (irregular) { pl->irregular = 1; continue; }
        if (us_b) { atomicAdd(&pl->n_usage_b, us_b); atomicMax(&pl->last_usage_pos, last_usage); }
        if (a_usage) pl->a_usage = 1;
    }

    // (3) post the event counters: one atomic per warp when the whole warp worked on one segment
    {
        const uint32_t full = 0xFFFFFFFFu;
        const uint32_t seg0 = __shfl_sync(full, acc_seg, 0);
        const bool uniform = __all_sync(full, acc_seg == seg0);
        if (uniform) {
            if (seg0 != 0xFFFFFFFFu) {
                const uint32_t sa = __reduce_add_sync(full, ev_a), sb = __reduce_add_sync(full, ev_b);
                if ((tid & 31) == 0) { if (sa) atomicAdd(&a.s.plan[seg0].n_events_a, sa); if (sb) atomicAdd(&a.s.plan[seg0].n_events_b, sb); }
            }
        } else if (acc_seg != 0xFFFFFFFFu) {
            if (ev_a) atomicAdd(&a.s.plan[acc_seg].n_events_a, ev_a);
            if (ev_b) atomicAdd(&a.s.plan[acc_seg].n_events_b, ev_b);
        }
    }
}

// ---- k_commit --------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64) k_commit(StepArgs a) {
    const uint32_t seg = blockIdx.x * blockDim.x + threadIdx.x;
    if (seg >= a.n_segs) return;
    const uint32_t c1 = a.seg_chunk[seg + 1];
    const uint32_t slot = a.seg_slot[seg];
    StreamHdr st = a.t.state[slot].h;
    const StepIO io = make_io(a, slot, &st);
    const SegPlan p = a.s.plan[seg];
    uint32_t emit_begin = p.emit_chunk_begin;
    if (st.phase == PH_COMMITTED && p.resume_chunk < c1) {
        const uint8_t* __restrict__ d = a.data;
        bool sequential = p.irregular || p.n_usage_b > 1;      // several usage candidates: let the exact path count them
        UsageRaw raw; uint32_t uf = 0;
        if (!sequential && p.last_usage_pos) {                  // validate the candidate with the full machine
            const uint32_t ps = p.last_usage_pos - 1;
            uint32_t pe = ps;
            while (pe + 1 < p.seg_end && !(__ldg(d + pe) == '\n' && __ldg(d + pe + 1) == '\n')) ++pe;
            Rope r{nullptr, 0, d + ps, pe - ps};
            const uint8_t cls = classify_part(r, 0, pe - ps);
            uf = parse_part<true>(r, 0, pe - ps, cls, &raw);
            if ((uf & PF_EXOTIC) || ((uf & TK_CHOICES) && (uf & PF_TYPE_ERROR))) sequential = true;
        }
        if (sequential) {
            run_chunks(io, a.data, a.chunk_off, p.resume_chunk, c1, emit_begin, false);
        } else {
            const uint32_t nch = c1 - p.resume_chunk, nby = p.seg_end - p.relay_begin;
            st.n_chunks_in += nch; st.n_chunks_emitted += nch; st.bytes_in += nby; st.bytes_emitted += nby;
            st.n_events_a += p.n_events_a; st.n_events_b += p.n_events_b; st.n_usage_b += p.n_usage_b;
            if (p.a_usage) st.flags |= SF_A_USAGE_BOUND;
            if (p.last_usage_pos) {                 // the last usage-bearing event wins (chat_logging.py:134-135)
                normalise_usage(raw, uf, *io.rec);
                st.flags |= SF_REC_VALID;
                if (io.rec->exotic) { ++st.n_exotic; st.flags |= SF_EXOTIC_SEEN; }
            }
            // new carry = text after the last separator (both loops: SF_SYNCED)
            uint32_t tail = p.relay_begin;
            if (nby >= 2 && __ldg(d + p.seg_end - 1) == '\n' && __ldg(d + p.seg_end - 2) == '\n') tail = p.seg_end;
            else {
                const uint32_t limit = nby > a.t.carry_cap + 2 ? p.seg_end - a.t.carry_cap - 2 : p.relay_begin;
                for (uint32_t k = p.seg_end; k >= limit + 2; --k)
                    if (__ldg(d + k - 1) == '\n' && __ldg(d + k - 2) == '\n') { tail = k; break; }
            }
            const uint32_t n = p.seg_end - tail;
            if (n > a.t.carry_cap) { st.carry_a_len = 0; st.flags |= SF_CARRY_OVERFLOW; }
            else { for (uint32_t k = 0; k < n; ++k) io.carry_a[k] = d[tail + k]; st.carry_a_len = n; }
        }
        a.t.state[slot].h = st;
    }
    SegResult res;
    fill_seg_result(st, emit_begin, c1, res);
    a.seg_out[seg] = res;
}

static inline cudaError_t launch_step_fast(const StepArgs& a, int sm_count, cudaStream_t stream, cudaEvent_t* ev, int* launched) {
    (void)sm_count;
    cudaError_t r;
    const uint32_t n_tiles = (a.n_bytes + LGW_TILE_BYTES - 1) / LGW_TILE_BYTES;
    const uint32_t n_prime = (a.n_segs > n_tiles + 1 ? a.n_segs : n_tiles + 1);
    k_prime<<<(n_prime + 63) / 64, 64, 0, stream>>>(a, n_tiles); ++*launched;
    if ((r = cudaEventRecord(ev[1], stream)) != cudaSuccess) return r;
    if (n_tiles) { k_relay<<<n_tiles, LGW_RELAY_THREADS, 0, stream>>>(a); ++*launched; }
    if ((r = cudaEventRecord(ev[2], stream)) != cudaSuccess) return r;
    if (a.n_segs) { k_commit<<<(a.n_segs + 63) / 64, 64, 0, stream>>>(a); ++*launched; }
    if ((r = cudaEventRecord(ev[3], stream)) != cudaSuccess) return r;
    return cudaGetLastError();
}
