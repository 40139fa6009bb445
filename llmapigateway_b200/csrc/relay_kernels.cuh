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
// vector store stays one STS.128 (the four words are permuted inside their own vector):
//   physical offset = d ^ X(d),  X(d) = ((d >> 3) & 0x30) | ((d >> 7) & 0x0c)
__device__ __forceinline__ uint32_t swz(uint32_t d) { return d ^ (((d >> 3) & 0x30u) | ((d >> 7) & 0x0cu)); }

// Shared memory of k_relay, declared at file scope so that device functions index it by name
// (LDS with an immediate base; a pointer would be rebuilt from the CTA's shared window per access).
__shared__ __align__(16) uint8_t sh_tile[LGW_TILE_BYTES];
__shared__ __align__(4) uint8_t sh_cls[256];
__shared__ __align__(4) uint8_t sh_trans[LGW_LEAN_ROWS * 32];
__shared__ uint32_t sh_seg_lo, sh_seg_hi;

// shared-memory loads by 32-bit shared-window address held in a register (the compiler otherwise
// rebuilds the CTA's shared-window base -- S2UR SR_CgaCtaId + ULEA -- in front of every access)
__device__ __forceinline__ uint32_t lds_u8(uint32_t a) { uint32_t v; asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { uint32_t v; asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t opaque(uint32_t x) { asm volatile("" : "+r"(x)); return x; }

// byte source + tables of the bulk kernel: the staged tile, global memory outside it
struct TileEnv {
    uint32_t tile_s, cls_s, trans_s;   // shared-window addresses (opaque registers)
    const uint8_t* __restrict__ g;     // whole packed buffer
    uint32_t t0, n_bytes;
    // aligned 32-bit word containing byte `pos` (little endian); bytes past n_bytes read as 0
    __device__ __forceinline__ uint32_t word(uint32_t pos) const {
        const uint32_t p4 = pos & ~3u, d = p4 - t0;
        if (d < LGW_TILE_BYTES) return lds_u32(tile_s + swz(d));
        return word_global(p4);
    }
    __device__ __noinline__ uint32_t word_global(uint32_t p4) const {
        if (p4 + 4 <= n_bytes) return __ldg(reinterpret_cast<const uint32_t*>(g + p4));
        uint32_t w = 0;
        for (uint32_t k = 0; k < 4 && p4 + k < n_bytes; ++k) w |= (uint32_t)__ldg(g + p4 + k) << (8 * k);
        return w;
    }
    __device__ __forceinline__ uint32_t at(uint32_t pos) const { return (word(pos) >> (8 * (pos & 3u))) & 0xffu; }
    __device__ __forceinline__ uint32_t cls(uint32_t c) const { return lds_u8(cls_s + c); }
    __device__ __forceinline__ uint32_t trans(uint32_t i) const { return lds_u8(trans_s + i); }
};

// rare path: chunk-level UTF-8 validation (request_handler.py:111 decodes each chunk on its own)
__device__ __noinline__ bool chunk_utf8_ok(const TileEnv* rd, uint32_t o, uint32_t e) {
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
__device__ __noinline__ bool find_open_event_start(const TileEnv* rd, uint32_t o, uint32_t relay_begin, uint32_t seg_end, uint32_t carry_cap, uint32_t* out_b) {
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
    const uint32_t t0 = blockIdx.x * LGW_TILE_BYTES;
    const uint32_t tid = threadIdx.x;
    const uint32_t n_bytes = a.n_bytes;
    const uint32_t c_lo = a.s.tile_chunk[blockIdx.x], c_hi = a.s.tile_chunk[blockIdx.x + 1];

    // (0) tables + segment range of this tile's chunks
    if (tid < 64) reinterpret_cast<uint32_t*>(sh_cls)[tid] = reinterpret_cast<const uint32_t*>(g_lean_tables_dev.cls)[tid];
    else if (tid < 64 + LGW_LEAN_ROWS * 8) reinterpret_cast<uint32_t*>(sh_trans)[tid - 64] = reinterpret_cast<const uint32_t*>(g_lean_tables_dev.trans)[tid - 64];
    if (tid >= 224 && tid < 226 && c_hi > c_lo) {
        const uint32_t c = tid == 224 ? c_lo : c_hi - 1;
        uint32_t lo = 0, hi = a.n_segs;
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (__ldg(a.seg_chunk + mid + 1) <= c) lo = mid + 1; else hi = mid; }
        if (tid == 224) sh_seg_lo = lo; else sh_seg_hi = lo;
    }

    // (1) re-emit: position-preserving 16-byte copy of the tile, staged (swizzled) into shared memory
    //     on the way; note whether the tile has any byte >= 0x80 (UTF-8 checks are skipped otherwise)
    uint32_t high = 0;
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
                uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0;
                for (uint32_t b = pos; b < n_bytes; ++b) {
                    const uint32_t c = a.data[b]; a.out[b] = (uint8_t)c;
                    const uint32_t sh = c << (8 * ((b - pos) & 3)), wi = (b - pos) >> 2;
                    if (wi == 0) w0 |= sh; else if (wi == 1) w1 |= sh; else if (wi == 2) w2 |= sh; else w3 |= sh;
                }
                x[k] = make_uint4(w0, w1, w2, w3);
            }
            high |= x[k].x | x[k].y | x[k].z | x[k].w;
            const uint32_t r = v >> 2;                          // 64-byte row
            const uint32_t kx = (r >> 3) & 3u;                  // word permutation inside the vector
            uint4 y;
            y.x = kx == 0 ? x[k].x : kx == 1 ? x[k].y : kx == 2 ? x[k].z : x[k].w;
            y.y = kx == 0 ? x[k].y : kx == 1 ? x[k].x : kx == 2 ? x[k].w : x[k].z;
            y.z = kx == 0 ? x[k].z : kx == 1 ? x[k].w : kx == 2 ? x[k].x : x[k].y;
            y.w = kx == 0 ? x[k].w : kx == 1 ? x[k].z : kx == 2 ? x[k].y : x[k].x;
            const uint32_t slot16 = (r << 6) | (((v & 3u) ^ ((r >> 1) & 3u)) << 4);
            *reinterpret_cast<uint4*>(sh_tile + slot16) = y;
        }
    }
    const int tile_high = __syncthreads_or((high & 0x80808080u) != 0);

    // (2) events of the chunks that START in this tile
    TileEnv env;
    env.tile_s = opaque((uint32_t)__cvta_generic_to_shared(sh_tile));
    env.cls_s = opaque((uint32_t)__cvta_generic_to_shared(sh_cls));
    env.trans_s = opaque((uint32_t)__cvta_generic_to_shared(sh_trans));
    env.g = a.data; env.t0 = t0; env.n_bytes = n_bytes;
    const uint32_t seg_lo = sh_seg_lo, seg_hi = sh_seg_hi;
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

        // chunk-level UTF-8: ASCII tiles need no check for chunks that end inside the tile
        if ((tile_high || e > t0 + LGW_TILE_BYTES) && !chunk_utf8_ok(&env, o, e)) { pl->irregular = 1; continue; }

        // where does the event that is open at the start of this chunk begin?
        uint32_t b = o;
        if (o != relay_begin) {
            const bool sep_before = o >= relay_begin + 2 && env.at(o - 1) == '\n' && env.at(o - 2) == '\n';
            if (sep_before) {
                if ((o >= relay_begin + 3 && env.at(o - 3) == '\n') || env.at(o) == '\n') { pl->irregular = 1; continue; }   // LF run >= 3
            } else if (!find_open_event_start(&env, o, relay_begin, seg_end, a.t.carry_cap, &b)) { pl->irregular = 1; continue; }
        }

        // walk the events that complete inside this chunk (second LF of the separator in [o, e))
        bool irregular = false;
        uint32_t us_b = 0, a_usage = 0, last_usage = 0;
        uint32_t ps = b;
        while (ps < e) {
            // classify the event prefix: "data: {" (handler + tap), "{" (tap only), anything else is skipped
            uint32_t cls = PC_NONE;
            {
                const uint32_t w0 = env.word(ps) >> (8 * (ps & 3u));
                if ((w0 & 0xffu) == '{') cls = PC_BRACE;
                else if ((w0 & 0xffu) == 'd') {
                    if ((ps & 3u) == 0) cls = (w0 == 0x61746164u && (env.word(ps + 4) & 0xFFFFFFu) == 0x7b203au) ? PC_DATA : PC_NONE;
                    else cls = (env.at(ps + 1) == 'a' && env.at(ps + 2) == 't' && env.at(ps + 3) == 'a' && env.at(ps + 4) == ':' && env.at(ps + 5) == ' ' && env.at(ps + 6) == '{') ? PC_DATA : PC_NONE;
                }
            }
            LeanMachine lm;
            lm.reset(cls == PC_DATA);
            uint32_t pos = ps + (cls == PC_DATA ? 6u : 0u);      // "data: " holds no LF
            bool ended = false;
            while (pos < e && !ended) {                          // word-wise byte loop
                uint32_t w = env.word(pos) >> (8 * (pos & 3u));
                uint32_t nb = 4 - (pos & 3u);
                if (nb > e - pos) nb = e - pos;
#pragma unroll 1
                for (; nb; --nb, w >>= 8, ++pos) {
                    const uint32_t ch = w & 0xffu;
                    // hot path: a plain byte inside a string changes nothing
                    if (lm.st == L_STR && ch >= 0x20u && ch != '"' && ch != '\\') continue;
                    if (ch == '\n') {
                        if (pos + 1 >= e) { pos = e; break; }     // a separator starting on the last byte completes later
                        if (env.at(pos + 1) == '\n') { ended = true; break; }
                    }
                    if (cls != PC_NONE) lm.step(ch, pos, env);
                }
            }
            if (!ended) break;                                   // the open event completes in a later chunk
            // ---- one complete event [ps, pos) ----
            if (pos + 2 < seg_end && env.at(pos + 2) == '\n') { irregular = true; break; }   // LF run >= 3
            if (cls != PC_NONE) {
                const uint32_t f = lm.finish();
                if (cls == PC_DATA) {                            // handler loop, request_handler.py:122-134
                    ++ev_a;
                    if ((f & PF_VALID_A) && !(f & TK_CODE) && (f & TK_USAGE)) a_usage = 1;
                }
                if (f & PF_VALID_B) {                            // tap loop, chat_logging.py:123-141
                    ++ev_b;
                    // events with "error" (extra DB row) go to the sequential path; a "usage" event is
                    // a candidate that k_commit validates with the full machine (choices walk)
                    if (f & TK_ERROR) { irregular = true; break; }
                    if (f & TK_USAGE) { ++us_b; last_usage = ps + 1; }
                }
            }
            ps = pos + 2;
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
        uint32_t ups = 0, upe = 0;
        if (!sequential && p.last_usage_pos) {                  // the winning usage event: [ups, upe)
            ups = p.last_usage_pos - 1; upe = ups;
            while (upe + 1 < p.seg_end && !(__ldg(d + upe) == '\n' && __ldg(d + upe + 1) == '\n')) ++upe;
            if (upe - ups > LGW_PENDING_CAP) sequential = true;
        }
        if (sequential) {
            run_chunks(io, a.data, a.chunk_off, p.resume_chunk, c1, emit_begin, false);
        } else {
            const uint32_t nch = c1 - p.resume_chunk, nby = p.seg_end - p.relay_begin;
            st.n_chunks_in += nch; st.n_chunks_emitted += nch; st.bytes_in += nby; st.bytes_emitted += nby;
            st.n_events_a += p.n_events_a; st.n_events_b += p.n_events_b;
            if (p.a_usage) st.flags |= SF_A_USAGE_BOUND;
            if (p.last_usage_pos) {                 // the last usage-bearing event wins (chat_logging.py:134-135):
                if (st.flags & SF_PENDING) resolve_pending(io);     // an older stash must be settled first
                for (uint32_t k = 0; k < upe - ups; ++k) io.pending[k] = __ldg(d + ups + k);    // stash its text;
                st.pending_len = upe - ups; st.flags |= SF_PENDING; ++st.n_usage_b;               // values on demand
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
