// swap.h — the virtual-device-memory swap engine (one per device/context).
//
// Behind the same switch as the reference's oversubscription (CUDA_OVERSUBSCRIBE=true -> allocmode 0,
// postInit libvgpu.so@0x15d63; large allocations then go through cuMemoryAllocate@0x315da), but where the
// reference hands the problem to CUDA UVM (cuMemAllocManaged; victim choice and page traffic inside the NVIDIA
// kernel driver on host cores), this engine keeps explicit control:
//   * every swappable allocation is a stable virtual range (cuMemAddressReserve arena) backed by VMM physical
//     handles that are mapped only while the buffer is resident;
//   * a device-resident allocation table (32-byte rows) records residency and a logical LRU clock; a GPU victim scan
//     picks exact-LRU victims from it;
//   * ONE PAGER THREAD per engine owns every driver call that is slow or that feeds the host link: cuMemUnmap /
//     cuMemMap / cuMemSetAccess / cuMemCreate (batched: one cuMemSetAccess / cuMemUnmap per run of adjacent ranges),
//     the victim scan, and the enqueueing of both DMA directions. Application threads (kernel launches, memcpys,
//     allocations) only do bookkeeping under a mutex, hand their misses to the pager and order their stream behind
//     the page-in's event. A stall inside the driver's VMM calls (0.1-20 ms under load on B200 / driver 580,
//     profiles/README.md) therefore no longer idles the link: the queues it feeds are deep, and it is not the
//     application thread that sits in the driver;
//   * two data paths between HBM and pinned host memory:
//       direct  — cuMemcpyDtoHAsync / HtoDAsync straight between the buffer's own range and its pinned block, in
//                 copy_bytes pieces: no staging, no kernel, no extra HBM traffic. Used whenever the pipeline runs ahead
//                 of need (prefetch + eviction ahead), i.e. for the bulk of the traffic;
//       staged  — the TMA pack kernel compacts the victims into a staging ring at HBM speed (their physical memory is
//                 free for the incoming buffer ~25 us later instead of one PCIe transfer later) while the copy
//                 stream drains the ring; page-in is the mirror image (pinned -> ring -> unpack kernel). Used for a
//                 demand miss that finds neither free physical memory nor an eviction in flight: latency path;
//   * a successor predictor (for every row: which row was touched next last time) drives prefetch: training loops and
//     sweeps repeat their access sequence, so the pager pages the next rows in — and evicts LRU rows ahead for them —
//     before the application asks; it only acts while its recent predictions were right (random access: off);
//   * clean rows: a row that was paged in and not written since keeps its pinned block; evicting it is an unmap, no
//     copy. Kernel launches count as writes unless the range was advised read-mostly (cuMemAdvise SET_READ_MOSTLY,
//     the hint UVM applications already give); memcpy sources never dirty.
#pragma once
#include <cuda.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "kernels.h"
#include "kmod.h"
#include "vgpu_region.h"

namespace vgpu {

struct SwapConfig {
    uint64_t resident_cap = 0;            // bytes of physical backing the container may hold (its gpumem quota); 0 = device free memory
    uint64_t virtual_cap = 0;             // live swappable bytes allowed; 0 = bounded by host pool only
    uint64_t host_pool_cap = 0;           // pinned bytes allowed; 0 = unbounded
    size_t chunk_bytes = 32u << 20;       // staging slot size (staged path: one DMA per slot)
    int ring_slots = 4;                   // staging slots per direction
    size_t slab_bytes = 1ull << 30;       // pinned pool growth unit
    uint64_t arena_bytes = 1ull << 40;    // virtual address arena
    bool profile = false;                 // bracket pack/unpack launches with events + in-kernel spans (bench roofline pass)
    uint32_t scan_lookahead = 16;         // a scan selects this many times the bytes needed; the surplus is consumed by later evictions
    uint64_t prefetch_bytes = 2048ull << 20;  // VGPU_SWAP_PREFETCH_MB: how far the pager runs ahead of the application, capped at a quarter
                                          // of the resident cap (0 = no prefetch). Deep on purpose: VMM calls stall for 10-500 ms now and
                                          // then while DMA is running (profiles/r02_*), and only queued work keeps the link busy meanwhile
    size_t copy_bytes = 16u << 20;        // VGPU_SWAP_COPY_MB: piece size of the direct copies (VMM calls wait for the copy in flight)
    uint32_t batch_rows = 4;              // rows per pager batch (VMM cost is per mapping, not per call: small batches keep latency low)
    uint64_t headroom_bytes = ~0ull;      // VGPU_SWAP_HEADROOM_MB: free physical memory the pager keeps ahead of the page-in queue while
                                          // the prefetch pipeline runs (~0 = as much as the prefetch window)
    bool host_backed = false;             // VGPU_SWAP_HOST_BACKED=1: an evicted range is re-mapped onto its host copy (VMM host
                                          // memory) instead of being left unmapped, so an access the hook could not see is slow, not fatal
    static SwapConfig from_env(uint64_t resident_cap, uint64_t virtual_cap);
};

struct SwapStats {
    uint64_t page_out_bytes = 0, page_in_bytes = 0;
    uint64_t evictions = 0, faults = 0, admissions = 0;
    uint64_t pack_launches = 0, unpack_launches = 0, scan_launches = 0, scans = 0, scan_cache_hits = 0;
    uint64_t resident_bytes = 0, live_bytes = 0, host_bytes = 0, entries = 0;
    uint64_t phys_creates = 0, phys_reuses = 0;
    uint64_t host_slabs = 0, host_slabs_local = 0;   // pinned slabs allocated / of those on the GPU's NUMA node
    double pack_ms = 0, unpack_ms = 0;     // CUDA-event brackets around each launch, only when profiling
    double pack_span_ms = 0, unpack_span_ms = 0;   // exact execution spans from in-kernel %globaltimer stamps (profiling)
    // APPLICATION-thread time inside ensure_resident/alloc (ns): whole admissions, and the part of it spent blocked until
    // the pager had issued the page-in. host_vmm_ns is the time application threads spend in VMM calls: zero by
    // construction since the pager owns them (kept in the ABI so that a regression shows).
    uint64_t host_admit_ns = 0, host_wait_ns = 0, host_vmm_ns = 0;
    // pager-thread time (ns): VMM calls, victim scans incl. their sync, waiting for the last pack of a staged batch,
    // staging-ring back-pressure
    uint64_t pager_vmm_ns = 0, pager_scan_ns = 0, pager_packsync_ns = 0, pager_ring_ns = 0, pager_busy_ns = 0;
    uint64_t vmm_calls = 0;                // cuMemUnmap + cuMemSetAccess calls issued (after batching)
    uint64_t pager_unmap_ns = 0, pager_setaccess_ns = 0, pager_map_ns = 0, pager_create_ns = 0;   // breakdown of pager_vmm_ns (diagnostics)
    uint64_t inplace_uses = 0;             // host-backed mode: operands of oversized launches used in place (host-mapped)
    uint64_t vmm_slow_calls = 0, vmm_slow_ns = 0, vmm_max_ns = 0;   // VMM calls that took > 2 ms (driver stalls), their time, the worst one
    uint64_t pager_issue_ns = 0, pager_poll_ns = 0, pager_lock_ns = 0, pager_step_ns[5] = {0, 0, 0, 0, 0};   // copy/event enqueue calls; busy time per step (zombies, reap, demand, prefetch, evict-ahead)
    uint64_t pack_bytes = 0, unpack_bytes = 0;       // bytes moved by the staged path's kernels
    uint64_t direct_out_bytes = 0, direct_in_bytes = 0;   // bytes moved by the direct path (subset of page_out/in_bytes)
    uint64_t prefetch_issued = 0, prefetch_hits = 0, prefetch_wasted = 0;   // rows paged in ahead / touched afterwards / evicted untouched
    uint64_t demand_waits = 0;             // admissions that had to block for the pager
    uint64_t clean_evictions = 0;          // evictions that needed no copy (host block still valid)
};

class SwapEngine {
   public:
    // Must be called with the target context current. Returns nullptr (and logs) on failure.
    static SwapEngine *create(int dev, const SwapConfig &cfg);
    ~SwapEngine();

    CUresult alloc(CUdeviceptr *dptr, size_t bytes);
    CUresult free(CUdeviceptr dptr);            // CUDA_ERROR_INVALID_VALUE when not one of ours
    bool owns(CUdeviceptr p) const { return p >= arena_ && p < arena_ + cfg_.arena_bytes; }
    int lookup(CUdeviceptr p) const;            // row index of the allocation containing p, or -1
    bool range_of(CUdeviceptr p, CUdeviceptr *base, size_t *size) const;
    // can these rows be resident at the same time (ensure_resident would not refuse them for their size)? no side effects
    bool fits_together(const int *rows, int n) const;   // the allocation containing p as the application made it

    // Makes every listed row resident and orders `stream` after the page-ins. Rows stay pinned (not evictable)
    // until note_use() is called with the same list after the real launch has been enqueued.
    // stream == kHostWait: block the calling thread until the rows are usable instead of ordering a stream behind the
    // page-ins (stream capture: a capturing stream may not wait on outside work). Rows acquired this way and never
    // passed to note_use() stay pinned resident — what a captured kernel's operands need.
    static inline CUstream kHostWait = reinterpret_cast<CUstream>(~uintptr_t(0));
    CUresult ensure_resident(const int *rows, int n, CUstream stream);
    // writes = false: the work just enqueued only READS the rows (memcpy source): they stay clean
    // closes_admission = false: a second call for the same admission (its read-only operands after its written ones)
    void note_use(const int *rows, int n, CUstream stream, bool writes = true, bool closes_admission = true);
    // Scans kernel parameter bytes for pointers into the arena; appends distinct row indices.
    void collect_rows(const void *param, size_t bytes, std::vector<int> *rows) const;
    // cuMemAdvise(SET/UNSET_READ_MOSTLY) on a swappable range: kernel launches no longer mark it dirty
    void advise_read_mostly(int row, bool on);
    // cuMemPrefetchAsync(range, device): queue a page-in with the pager, do not wait
    void hint_prefetch(int row);
    void hint_evict(int row);                   // cuMemPrefetchAsync(range, CU_DEVICE_CPU): first in line when room is needed
    // never evict this row (operands the argument scan cannot see: device-side pointer tables)
    CUresult pin_resident(int row, bool on);

    SwapStats stats();
    void set_profile(bool on);
    // the quota left for swappable memory shrinks/grows with the container's non-swappable bytes (context, small buffers)
    void set_resident_cap(uint64_t cap);
    // publish the counters into the container's shared region (vgpu_region.h extension block) after every call that
    // changes them; nullptr = do not publish
    // Several processes of one container on one device share ONE resident quota: before the pager lets the resident set
    // grow it asks this function (the hook implements it on the container's shared region, under its lock) whether
    // `want_total` resident bytes fit next to what the sibling engines hold or are entitled to; the answer is also this
    // engine's current cap. Without it (one engine per device, the C ABI) the cap is the one given at creation.
    typedef std::function<uint64_t(uint64_t want_total, uint64_t live_mapped, bool *granted, int *engines, uint64_t *fair_share)> BudgetFn;
    void set_budget_fn(BudgetFn fn) { std::lock_guard<std::mutex> g(mu_); budget_fn_ = std::move(fn); kick_pager_locked(); }
    void set_shared_record(vgpu_swap_record_t *rec) { std::lock_guard<std::mutex> g(mu_); shared_ = rec; publish_locked(); }
    CUresult drain();                          // wait for the pager and all side-stream work (tests / shutdown)
    void dump_trace(FILE *f);                  // VGPU_SWAP_TRACE=n: device-side start/end times of the first n direct copies per direction after warm-up
    void stop_pager();                         // process exit: no driver call may be in flight when the driver deinitialises
    const SwapConfig &config() const { return cfg_; }
    // device memory the engine itself holds next to the application's resident buffers: both staging rings (the table,
    // scan scratch and span words are a few hundred KiB and not counted). The hook takes it out of the room it gives
    // the engine, so that the container's PHYSICAL footprint stays within its gpumem quota.
    uint64_t device_overhead() const { return 2ull * cfg_.ring_slots * cfg_.chunk_bytes; }
    uint64_t live_bytes() const { return live_bytes_.load(std::memory_order_relaxed); }
    uint64_t resident_bytes() const { return resident_pub_.load(std::memory_order_relaxed); }

    // test hook: copy of the host mirror of the table
    std::vector<VgpuEntry> snapshot_table();

   private:
    enum Phase : uint8_t {
        PH_IDLE = 0,      // stable: resident or paged out
        PH_QUEUED,        // paged out, waiting in demand_q_/prefetch_q_
        PH_LOADING,       // the pager is mapping it / enqueueing its page-in (mu_ released meanwhile)
        PH_EVICTING,      // page-out enqueued, still mapped; logically paged out
        PH_ZOMBIE,        // freed by the application while resident: the pager unmaps it once its last users are done
    };
    static constexpr int kMaxUses = 4;
    struct Side {                // host-only companion of a table row
        size_t mapped = 0;
        CUmemGenericAllocationHandle handle = 0;
        bool has_handle = false;
        uint64_t host_off = 0;   // pinned block (kept across page-ins: a clean row is evicted without a copy)
        bool has_host = false;
        uint64_t zombie_len = 0; // free(): the block's length, kept for the pager (the table row's size is zeroed at once)
        bool dirty = false;      // HBM content differs from the pinned block (or there is none yet and the row was written)
        bool read_mostly = false;
        bool prefetched = false; // paged in ahead of need and not touched since
        bool demand = false;     // an application thread is waiting for this row
        bool locked = false;     // pin_resident(): never a victim
        Phase phase = PH_IDLE;
        CUresult fail = CUDA_SUCCESS;   // why the pager could not bring it in (reported to the waiting admission)
        CUevent ready = nullptr; // pending page-in completion
        uint64_t uses[kMaxUses] = {0, 0, 0, 0};   // sequence numbers of the outstanding last-use events, one per stream
        int nuses = 0;
        int pins = 0;
        int inplace = 0;         // host-backed mode: admissions that were told to use the row where it is (host-mapped); no page-in meanwhile
        uint64_t va_off = 0;
        CUevent evict_done = nullptr;   // after it the row's physical memory is no longer read by its page-out
        int out_slot = -1;       // staging slot of the last staged page-out chunk of this row ...
        uint64_t out_seq = 0;    // ... and that slot's use counter at the time (stale => the D2H is known complete)
        // host-backed mode (VGPU_SWAP_HOST_BACKED=1): the pinned block is a host-located VMM handle of its own, mapped for good at
        // the row's alias address (harena_ + va_off, the DMA target) and, WHILE THE ROW IS PAGED OUT, also at the row's own
        // address — an access the hook could not see (a pointer table in device memory) then reads and writes host memory
        // over PCIe instead of faulting on an unmapped range, which is what UVM gives the reference
        CUmemGenericAllocationHandle hhandle = 0;
        bool has_hh = false;
        bool hosted = false;     // the row's own range currently maps hhandle
        uint32_t gen = 0;        // bumped when the row index is recycled (stale queue entries are skipped)
        int retries = 0;         // page-in attempts that met a device with less memory than the cap promises
    };
    struct Slab { unsigned char *host = nullptr; size_t bytes = 0; std::map<uint64_t, uint64_t> free; };
    struct Slot { CUdeviceptr buf = 0; CUevent busy = nullptr; bool used = false; uint64_t seq = 0; };
    typedef std::unique_lock<std::mutex> Lock;

    SwapEngine() = default;
    bool init(int dev, const SwapConfig &cfg);

    // ---- table (mu_)
    int new_row();
    void mark_dirty(int row);
    CUresult sync_table(Lock &lk, CUstream s);
    void publish_locked();
    bool row_live(int row) const { return row >= 0 && (size_t)row < rows_.size() && rows_[row].state != VGPU_ST_FREE; }
    void retire_row_locked(int row);          // give the index and the VA range back (row is unmapped)

    // ---- predictor (mu_)
    void observe_touch(int row);
    bool predictor_confident() const;
    void schedule_prefetch();

    // ---- pager thread
    void pager_main();
    bool step_zombies(Lock &lk);
    bool step_reap(Lock &lk);
    bool step_demand(Lock &lk);
    bool step_prefetch(Lock &lk);
    bool step_evict_ahead(Lock &lk);
    struct OutItem { uint32_t row; CUdeviceptr base; uint64_t len; size_t mapped; uint64_t host_off; bool has_host, copy; bool new_host = false; uint64_t va_off = 0; CUmemGenericAllocationHandle hh = 0; bool has_hh = false; std::vector<CUevent> wait; CUevent done = nullptr; int out_slot = -1; uint64_t out_seq = 0; bool failed = false; };
    struct InItem { int row; CUdeviceptr base; uint64_t len; size_t mapped; uint64_t host_off; bool has_host; bool prefetch; uint64_t va_off = 0; bool hosted = false; CUmemGenericAllocationHandle h = 0; CUevent ready = nullptr;
                    CUevent after = nullptr; CUresult rc = CUDA_SUCCESS; std::vector<CUevent> wait; };
    CUresult choose_victims(Lock &lk, uint64_t shortage, std::vector<uint32_t> *victims, uint64_t *evictable);
    void begin_evict_locked(const std::vector<uint32_t> &victims, std::vector<OutItem> *items);
    CUresult evict_direct(Lock &lk, const std::vector<uint32_t> &victims);
    void begin_load_locked(int row, bool prefetch, InItem *it);
    CUresult load_direct(Lock &lk, std::vector<InItem> &items);
    void commit_load_locked(InItem &it);
    void fail_load_locked(InItem &it, CUresult rc);
    bool requeue_under_pressure_locked(int row, bool was_demand, size_t free_dev);
    CUresult swap_staged(Lock &lk, int row, const std::vector<uint32_t> &victims);
    void unmap_batch(std::vector<std::pair<CUdeviceptr, size_t>> &ranges);
    CUresult make_host_handle(size_t mapped, uint64_t va_off, CUmemGenericAllocationHandle *h);
    void drop_host_handle(CUdeviceptr base, size_t mapped, uint64_t va_off, CUmemGenericAllocationHandle h, bool hosted);
    CUresult set_access_batch(std::vector<std::pair<CUdeviceptr, size_t>> &ranges);
    CUresult obtain_phys(size_t mapped, CUmemGenericAllocationHandle *h, bool *pressure);
    void pool_phys(size_t mapped, CUmemGenericAllocationHandle h);
    void trim_phys_pool(uint64_t need);
    int64_t free_phys_locked() const { return (int64_t)cfg_.resident_cap - (int64_t)resident_mapped_ - (int64_t)evicting_mapped_; }
    void flush_pager_stats_locked();
    void relock(Lock &lk);
    void note_vmm_call(const char *what, uint64_t ns);
    void collect_waits_locked(int row, std::vector<CUevent> *out);

    // ---- staged path pieces (pager thread)
    struct InRun { unsigned char *src; uint64_t pos, len; };
    struct InJob { Slot *slot = nullptr; std::vector<PackSegment> segs; std::vector<InRun> runs; uint64_t bytes = 0; };
    Slot &acquire_slot(std::vector<Slot> &ring, int *cursor);
    CUresult staged_out(std::vector<OutItem> &items, CUevent *packed);
    void plan_staged_in(const InItem &it, std::vector<InJob> *jobs);
    CUresult in_issue_copies(InJob &j);

    // ---- pinned pool (host_mu_)
    bool pin_slab(size_t sb, Slab *out, bool *local);
    void grow_main();
    void want_host_pool(uint64_t live_bytes);
    bool host_alloc(size_t bytes, uint64_t *off);
    void release_host_range(uint64_t off, uint64_t len);
    unsigned char *host_ptr(uint64_t off);
    bool reclaim_host_blocks(Lock &lk, uint64_t bytes);

    // ---- misc
    CUevent get_event();                       // shared pool (ev_mu_)
    void put_event(CUevent e);
    bool va_alloc(size_t bytes, uint64_t *off);
    void va_free(uint64_t off, size_t bytes);
    CUevent use_event(uint64_t seq);
    void prof_begin(CUstream s, CUevent *a);
    void prof_end(CUstream s, CUevent a, bool unpack, uint64_t bytes);
    void harvest_prof(bool wait);
    CUdeviceptr next_span(bool unpack);
    void harvest_spans();

    mutable std::mutex mu_;
    std::condition_variable cv_pager_;              // work for the pager
    std::condition_variable cv_admit_;              // progress for application threads (and drain)
    std::mutex gate_mu_;                            // one admission WITH misses at a time (a second one could hold the rows the first needs)
    int dev_ = 0;
    int numa_node_ = -1;                            // NUMA node of the GPU (pinned slabs are allocated there)
    SwapConfig cfg_;
    const Kernels *k_ = nullptr;
    size_t gran_ = 2u << 20;
    CUdeviceptr arena_ = 0;
    CUdeviceptr harena_ = 0;                        // host-backed mode: alias range, same size and layout as the arena
    CUcontext ctx_ = nullptr;
    bool ctx_warned_ = false;

    // ---- shared state (mu_)
    std::map<uint64_t, uint64_t> va_free_;          // offset -> len
    std::vector<int32_t> owner_;                    // granule -> row (or -1)
    std::vector<VgpuEntry> rows_;                   // host mirror (authoritative)
    std::vector<Side> side_;
    std::vector<int> free_rows_;
    uint32_t dirty_lo_ = UINT32_MAX, dirty_hi_ = 0;
    uint64_t tick_ = 0;
    uint64_t resident_mapped_ = 0;                  // mapped bytes of rows that are resident or loading
    uint64_t evicting_mapped_ = 0;                  // mapped bytes of rows whose eviction is in flight (still hold physical memory)
    std::atomic<uint64_t> live_bytes_{0}, resident_pub_{0};
    std::atomic<uint64_t> host_used_{0};
    struct QEntry { int row; uint32_t gen; };
    std::deque<QEntry> demand_q_, prefetch_q_;
    uint64_t queued_prefetch_bytes_ = 0, prefetched_bytes_ = 0;
    uint64_t live_mapped_ = 0;                      // mapped-size sum of the live rows (all resident => nothing to prefetch)
    std::deque<uint32_t> evicting_;                 // rows in PH_EVICTING, issue order
    std::deque<uint32_t> zombies_;
    std::vector<CUevent> use_ring_;                 // the last-use event recorded for sequence number seq, at seq % size
    std::vector<CUstream> use_stream_;              // stream each use event was recorded on ...
    std::vector<CUcontext> use_ctx_;                // ... and the context that stream belongs to (stream 0 exists once per context)
    // An event must be recorded on a stream of ITS OWN context, so every context the application uses the buffers from gets
    // its own set of last-use events (the engine context's set is own_events_). Waiting on them from the pager's streams and
    // on the page-in events from the application's streams works across contexts.
    std::vector<CUevent> own_events_;
    struct CtxEvents { CUcontext ctx; std::vector<CUevent> ev; };
    std::vector<CtxEvents> ctx_events_;
    CUevent use_slot_event(CUcontext cur, size_t slot);
    uint64_t use_seq_ = 0;
    uint64_t release_epoch_ = 0;                    // bumped by every note_use: pins were released
    int open_admissions_ = 0;                       // admissions whose use has not been recorded yet (ensure_resident .. note_use): their pins are about to go
    bool stop_ = false, pager_idle_ = true, kick_ = false;
    void kick_pager_locked() { kick_ = true; cv_pager_.notify_one(); }
    void drop_prefetch_queue_locked();
    void fail_demands_locked(CUresult rc);
    // predictor
    std::vector<int32_t> succ_;                     // row -> row touched right after it last time
    int last_row_ = -1;
    uint32_t pred_hist_ = 0, pred_count_ = 0;       // last 32 predictions (1 = right)
    SwapStats st_;
    BudgetFn budget_fn_;
    uint64_t budget_checked_ns_ = 0, fair_share_ = 0, demand_since_ns_ = 0;
    int demand_row_ = -1;
    int sibling_engines_ = 1;                       // engines of the container on this device (from the last budget query)
    bool reserve_locked(uint64_t extra);            // may the resident set grow by `extra`? (refreshes the cap from the shared budget)
    vgpu_swap_record_t *shared_ = nullptr;
    // Physical pressure: the quota (quota_cap_) promises more than the device can give right now — other containers of an
    // overcommitted GPU (DeviceMemoryScaling > 1, server.go:356) hold the rest. The working cap (cfg_.resident_cap) is
    // lowered to what was obtainable and probed back up every few admissions.
    uint64_t quota_cap_ = 0;
    bool pressure_ = false;
    uint32_t pressure_probe_ = 0;
    uint64_t pressure_events_ = 0;

    // ---- pager-private state (touched by the pager thread only, no lock)
    std::thread pager_;
    SwapStats pst_;                                 // deltas, folded into st_ under mu_
    std::multimap<size_t, CUmemGenericAllocationHandle> phys_pool_;
    uint64_t phys_pool_bytes_ = 0;
    CUdeviceptr d_tbl_ = 0;
    uint32_t tbl_cap_ = 0;
    VgpuEntry *h_tbl_stage_ = nullptr;              // pinned upload buffer
    CUdeviceptr dh_tbl_stage_ = 0;                  // ... and its device view (the upload is done by a kernel)
    // independent queues: a pack never waits behind an unpack that is itself waiting for its H2D, and a victim
    // scan never waits behind either
    CUstream s_scan_ = nullptr, s_pack_ = nullptr, s_unpack_ = nullptr, s_out_ = nullptr, s_in_ = nullptr;
    std::vector<Slot> ring_out_, ring_in_;
    int cur_out_ = 0, cur_in_ = 0;
    std::unique_ptr<VictimScanner> scanner_;
    std::mutex ev_mu_;
    std::vector<CUevent> ev_pool_;                  // (ev_mu_) recycled events: page-in / eviction completions, pack markers
    uint64_t scan_unit_ = 0;                        // mapped size of the last demanded row: the unit of the scan look-ahead
    uint32_t reap_defer_ = 0;                       // polls a partial reap batch has been held back
    bool unmap_runs_ok_ = true;                     // one cuMemUnmap may span several adjacent mappings (probed at run time)
    // victims selected by the last scan beyond what was needed then, in LRU order. They stay the exact LRU prefix for
    // as long as they are untouched (anything touched or created since carries a larger tick), so consuming them
    // is equivalent to re-scanning; an entry whose row changed is simply skipped.
    struct Cand { uint32_t row; uint64_t touch, base; };
    std::deque<Cand> victim_cache_;
    struct Prof { CUevent a, b; bool unpack; uint64_t bytes; };
    CUdeviceptr d_span_ = 0;                        // profiling: {min start, max end} per launch, pre-set to {~0, 0}
    uint32_t span_cap_ = 0, span_next_ = 0, span_read_ = 0;
    std::vector<uint8_t> span_unpack_;
    std::vector<Prof> prof_;
    std::atomic<bool> profile_{false};
    // diagnostics (VGPU_SWAP_TRACE): timed events around direct copies, host time of issue
    struct TraceRec { int dir; int row; uint64_t host_ns; CUevent a, b; };
    std::vector<TraceRec> trace_;
    uint32_t trace_want_ = 0, trace_skip_ = 0;
    CUevent trace_base_ = nullptr;
    uint64_t trace_base_ns_ = 0;
    bool trace_begin(int dir, int row, CUstream s);
    void trace_end(CUstream s);

    // ---- pinned pool (host_mu_)
    std::mutex host_mu_;
    std::condition_variable host_cv_;
    std::vector<Slab> slabs_;
    uint64_t host_total_ = 0, host_slabs_ = 0, host_slabs_local_ = 0;
    std::atomic<uint64_t> host_want_{0};            // pool size the grower works towards (= pinned blocks the live rows will need)
    uint64_t host_need_ = 0;                        // (mu_) sum of the live rows' block sizes
    std::thread grower_;
    bool growing_ = false, grow_stop_ = false;
};

}  // namespace vgpu
