// fsg_frame.cu — frame output: the legacy-VTK point cloud the reference drivers write every 10 / 20 steps through
// the LLNL VisIt writer (write_point_mesh, visit_writer.cpp:673-719; called at solver-unidyn.cu:487, commented out
// at solver.cu:213).  Same bytes as that writer for the same arrays, ASCII and binary — but re-entrant (no
// file-scope FILE*), buffered (the reference issues one fprintf per number), and it reports a file that cannot
// be opened instead of crashing (visit_writer.cpp:145 has no NULL check).
//
// File layout (visit_writer.cpp:327-335, 673-719, 358-644):
//   # vtk DataFile Version 2.0 / Written using VisIt writer / ASCII|BINARY / DATASET UNSTRUCTURED_GRID
//   POINTS n float      3n numbers, "%20.12e " each, 9 per line (ASCII) or big-endian floats (binary)
//   CELLS n 2n          "1 i" per point;   CELL_TYPES n  "1" (VISIT_VERTEX) per point
//   CELL_DATA n (empty);  POINT_DATA n;  first scalar as SCALARS <name> float / LOOKUP_TABLE default, first vector
//   as VECTORS <name> float, remaining scalars under FIELD FieldData k, then remaining vectors likewise.
#include "fsg_internal.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {
// Formatting threads for the big number arrays ("%20.12e " is 21 bytes for every float, so every number's place in the file is
// known up front and the ranges are formatted side by side).  FSG_FRAME_THREADS overrides (1 = serial).
int frame_threads()
{
    static int n = 0;
    if (!n) {
        const char *e = getenv("FSG_FRAME_THREADS");
        n = e ? atoi(e) : (int)std::thread::hardware_concurrency();
        if (n < 1) n = 1;
        if (n > 32) n = 32;
    }
    return n;
}
template <typename F>
void parallel_ranges(int64_t count, int64_t min_per_thread, F body)      // body(begin, end, part)
{
    int t = (int)std::min<int64_t>(frame_threads(), (count + min_per_thread - 1) / min_per_thread);
    if (t <= 1) { body((int64_t)0, count, 0); return; }
    std::vector<std::thread> th;
    const int64_t per = (count + t - 1) / t;
    for (int k = 0; k < t; k++) {
        const int64_t b = k * per, e = std::min(count, b + per);
        if (b >= e) break;
        th.emplace_back([=, &body] { body(b, e, k); });
    }
    for (auto &x : th) x.join();
}

struct VtkOut {
    FILE *fp = nullptr;
    bool binary = false;
    int col = 0;                 // numInColumn of the reference writer
    std::vector<char> buf;
    void flush() { if (!buf.empty()) { fwrite(buf.data(), 1, buf.size(), fp); buf.clear(); } }
    void raw(const void *p, size_t n) { const char *c = (const char *)p; buf.insert(buf.end(), c, c + n); if (buf.size() > (1u << 22)) flush(); }
    void str(const char *s) { raw(s, strlen(s)); }
    void end_line() { if (!binary) { raw("\n", 1); col = 0; } }                         // visit_writer.cpp:110-118
    void new_section() { if (col != 0) end_line(); col = 0; }                           // :236-241
    static void swap4(unsigned char *b) { unsigned char t = b[0]; b[0] = b[3]; b[3] = t; t = b[1]; b[1] = b[2]; b[2] = t; }
    void put_float(float v)                                                             // :295-312
    {
        if (binary) { unsigned char b[4]; memcpy(b, &v, 4); swap4(b); raw(b, 4); return; }
        char s[64];
        int n = snprintf(s, sizeof s, "%20.12e ", v);
        raw(s, (size_t)n);
        if (((col++) % 9) == 8) end_line();
    }
    // `count` floats in a row: the same bytes put_float would produce one by one
    void put_floats(const float *v, int64_t count)
    {
        if (count <= 0) return;
        flush();
        if (binary) {
            std::vector<unsigned char> out((size_t)count * 4);
            parallel_ranges(count, 1 << 18, [&](int64_t b, int64_t e, int) {
                for (int64_t j = b; j < e; j++) { unsigned char *p = &out[(size_t)j * 4]; memcpy(p, &v[j], 4); swap4(p); }
            });
            fwrite(out.data(), 1, out.size(), fp);
            return;
        }
        // element j starts at 21 j + (newlines before it); a newline follows element j when (col + j) % 9 == 8
        const int c0 = col;
        const int64_t total = 21 * count + (c0 + count) / 9;
        std::vector<char> out((size_t)total);
        bool ok = true;
        parallel_ranges(count, 1 << 15, [&](int64_t b, int64_t e, int) {
            char s[64];
            for (int64_t j = b; j < e; j++) {
                const int n = snprintf(s, sizeof s, "%20.12e ", v[j]);
                if (n != 21) { ok = false; return; }                      // (cannot happen for a float: two exponent digits at most)
                char *p = &out[(size_t)(21 * j + (c0 + j) / 9)];
                memcpy(p, s, 21);
                if ((c0 + j) % 9 == 8) p[21] = '\n';
            }
        });
        if (!ok) { for (int64_t j = 0; j < count; j++) put_float(v[j]); return; }
        fwrite(out.data(), 1, out.size(), fp);
        col = (int)((c0 + count) % 9);
    }
    // the CELLS section ("1 i" per point) and CELL_TYPES ("1" per point), formatted in parallel parts and written in order
    void put_cells(int npts)
    {
        flush();
        if (binary) { for (int i = 0; i < npts; i++) { put_int(1); put_int(i); } return; }
        std::vector<std::string> parts((size_t)frame_threads());
        parallel_ranges(npts, 1 << 15, [&](int64_t b, int64_t e, int part) {
            std::string &o = parts[(size_t)part];
            o.reserve((size_t)(e - b) * 12);
            char s[32];
            for (int64_t i = b; i < e; i++) { const int n = snprintf(s, sizeof s, "1 %d \n", (int)i); o.append(s, (size_t)n); }
        });
        for (auto &o : parts) if (!o.empty()) fwrite(o.data(), 1, o.size(), fp);
        col = 0;
    }
    void put_cell_types(int npts)
    {
        flush();
        if (binary) { for (int i = 0; i < npts; i++) put_int(1); return; }
        std::string o;
        o.reserve((size_t)npts * 3);
        for (int i = 0; i < npts; i++) o.append("1 \n", 3);
        fwrite(o.data(), 1, o.size(), fp);
        col = 0;
    }
    void put_int(int v)                                                                 // :254-275
    {
        if (binary) { unsigned char b[4]; memcpy(b, &v, 4); swap4(b); raw(b, 4); return; }
        char s[32];
        int n = snprintf(s, sizeof s, "%d ", v);
        raw(s, (size_t)n);
        if (((col++) % 9) == 8) { raw("\n", 1); col = 0; }
    }
};

// write_variables with centering == 1 for every variable (what write_point_mesh passes), :358-644
void put_point_variables(VtkOut &o, int nvars, const int *vardim, const char *const *names, const float *const *vars, int npts)
{
    char s[1024];
    o.new_section();
    snprintf(s, sizeof s, "CELL_DATA %d\n", npts);
    o.str(s);
    o.new_section();
    snprintf(s, sizeof s, "POINT_DATA %d\n", npts);
    o.str(s);
    int first_scalar = 0, first_vector = 0, num_scalars = 0, num_vectors = 0;
    for (int i = 0; i < nvars; i++) {
        bool write = false;
        if (vardim[i] == 1) {
            if (!first_scalar) { write = true; snprintf(s, sizeof s, "SCALARS %s float\n", names[i]); o.str(s); o.str("LOOKUP_TABLE default\n"); first_scalar = 1; }
            else num_scalars++;
        } else if (vardim[i] == 3) {
            if (!first_vector) { write = true; snprintf(s, sizeof s, "VECTORS %s float\n", names[i]); o.str(s); first_vector = 1; }
            else num_vectors++;
        } else
            continue;                         // the reference prints a warning and ignores the variable
        if (write) {
            o.put_floats(vars[i], (int64_t)npts * vardim[i]);
            o.end_line();
        }
    }
    for (int dim = 1; dim <= 3; dim += 2) {
        int count = dim == 1 ? num_scalars : num_vectors, first = 0;
        if (count <= 0) continue;
        snprintf(s, sizeof s, "FIELD FieldData %d\n", count);
        o.str(s);
        for (int i = 0; i < nvars; i++) {
            if (vardim[i] != dim) continue;
            if (!first) { first = 1; continue; }
            snprintf(s, sizeof s, "%s %d %d float\n", names[i], dim, npts);
            o.str(s);
            o.put_floats(vars[i], (int64_t)npts * dim);
            o.end_line();
        }
    }
}
}  // namespace

extern "C" int fsg_write_point_mesh(const char *filename, int use_binary, int npts, const float *pts, int nvars, const int *vardim,
                                    const char *const *varnames, const float *const *vars)
{
    if (!filename || npts < 0 || (npts > 0 && !pts) || nvars < 0 || (nvars > 0 && (!vardim || !varnames || !vars))) return FSG_E_INVALID;
    std::string full = filename;
    if (!strstr(filename, ".vtk")) full += ".vtk";                                       // open_file, :133-147
    VtkOut o;
    o.fp = fopen(full.c_str(), "w+");
    if (!o.fp) return FSG_E_INVALID;
    o.binary = use_binary != 0;
    o.str("# vtk DataFile Version 2.0\nWritten using VisIt writer\n");                   // write_header, :327-335
    o.str(o.binary ? "BINARY\n" : "ASCII\n");
    char s[128];
    o.str("DATASET UNSTRUCTURED_GRID\n");
    snprintf(s, sizeof s, "POINTS %d float\n", npts);
    o.str(s);
    o.put_floats(pts, 3ll * npts);
    o.new_section();
    snprintf(s, sizeof s, "CELLS %d %d\n", npts, 2 * npts);
    o.str(s);
    o.put_cells(npts);                                                                   // "1 i" + end of line per point
    o.new_section();
    snprintf(s, sizeof s, "CELL_TYPES %d\n", npts);
    o.str(s);
    o.put_cell_types(npts);                                                              // VISIT_VERTEX per point
    put_point_variables(o, nvars, vardim, varnames, vars, npts);
    o.end_line();                                                                        // close_file, :161-166
    o.flush();
    int rc = ferror(o.fp) ? FSG_E_INVALID : FSG_OK;
    fclose(o.fp);
    return rc;
}

// The frame a driver writes after a step: positions + the two scalars mykernel2 exported (solver.cu:108-109 "dens",
// "cellnumber"; solver-unidyn.cu:117 "mass", "surface_level").
extern "C" int fsg_write_frame(fsg_ctx *c, const char *filename, int use_binary)
{
    if (!c || !filename) return FSG_E_INVALID;
    const int64_t n = c->n;
    if (n > 0x7fffffff / 3) { c->err = "fsg_write_frame: too many particles for the legacy VTK point-cloud writer"; return FSG_E_INVALID; }
    std::vector<float> spts(3 * (size_t)n), a3((size_t)n), b3((size_t)n);
    int rc = fsg_export_viz(c, spts.data(), a3.data(), b3.data());
    if (rc != FSG_OK) return rc;
    const int vardim[2] = {1, 1};
    const bool uni = c->cfg.model == FSG_MODEL_UNIDYN;
    const char *names[2] = {uni ? "mass" : "dens", uni ? "surface_level" : "cellnumber"};
    const float *vars[2] = {a3.data(), b3.data()};
    rc = fsg_write_point_mesh(filename, use_binary, (int)n, spts.data(), 2, vardim, names, vars);
    if (rc != FSG_OK) c->err = std::string("fsg_write_frame: cannot write ") + filename;
    return rc;
}


// ------------------------------------------------------------------------------------------------
// Asynchronous frame output.  The reference writes a frame synchronously inside the time loop
// (solver-unidyn.cu:472-493: cudaDeviceSynchronize, managed arrays read by the host, one fprintf per number).  Here a frame costs
// the step loop ONE small export kernel (20 B per particle) on the context's stream; the device-to-host copy runs on a second
// stream behind an event, into one of two pinned staging slots, and a writer thread formats and writes the file while the
// solver keeps stepping.  Two slots: a third frame waits until the oldest one is on disk (back-pressure instead of unbounded
// memory).  fsg_frame_wait() joins the pending frames and reports the first error.
// ------------------------------------------------------------------------------------------------
struct FsgFrameWriter {
    struct Job { std::string filename; int binary; int64_t n; int slot; bool uni; };
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::deque<Job> q;
    bool stop = false, busy = false;
    bool slot_busy[2] = {false, false};
    int first_error = FSG_OK;
    std::string err;
    cudaStream_t copy = nullptr;
    cudaEvent_t exported[2] = {nullptr, nullptr}, copied[2] = {nullptr, nullptr};
    float *dev[2] = {nullptr, nullptr}, *host[2] = {nullptr, nullptr};
    int64_t cap[2] = {0, 0};
    int device = 0;
    int64_t frames_written = 0;

    void run()
    {
        cudaSetDevice(device);
        for (;;) {
            Job j;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return stop || !q.empty(); });
                if (q.empty()) return;
                j = q.front();
                q.pop_front();
                busy = true;
            }
            int rc = cudaEventSynchronize(copied[j.slot]) == cudaSuccess ? FSG_OK : FSG_E_CUDA;
            if (rc == FSG_OK) {
                const float *spts = host[j.slot], *a3 = spts + 3 * j.n, *b3 = a3 + j.n;
                const int vardim[2] = {1, 1};
                const char *names[2] = {j.uni ? "mass" : "dens", j.uni ? "surface_level" : "cellnumber"};
                const float *vars[2] = {a3, b3};
                rc = fsg_write_point_mesh(j.filename.c_str(), j.binary, (int)j.n, spts, 2, vardim, names, vars);
            }
            {
                std::lock_guard<std::mutex> lk(mu);
                if (rc != FSG_OK && first_error == FSG_OK) { first_error = rc; err = "frame output: cannot write " + j.filename; }
                if (rc == FSG_OK) frames_written++;
                slot_busy[j.slot] = false;
                busy = false;
            }
            cv.notify_all();
        }
    }
};

void fsg_frame_writer_destroy(fsg_ctx *c)
{
    FsgFrameWriter *w = c->frame_writer;
    if (!w) return;
    {
        std::lock_guard<std::mutex> lk(w->mu);
        w->stop = true;
    }
    w->cv.notify_all();
    if (w->th.joinable()) w->th.join();          // pending frames are written first (run() drains the queue before it returns)
    for (int k = 0; k < 2; k++) {
        if (w->dev[k]) cudaFree(w->dev[k]);
        if (w->host[k]) cudaFreeHost(w->host[k]);
        if (w->exported[k]) cudaEventDestroy(w->exported[k]);
        if (w->copied[k]) cudaEventDestroy(w->copied[k]);
    }
    if (w->copy) cudaStreamDestroy(w->copy);
    delete w;
    c->frame_writer = nullptr;
}

#define CUF(ctx, call)                                                                                  \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess) {                                                                        \
            (ctx)->err = std::string(#call) + " failed: " + cudaGetErrorString(e_);                     \
            return e_ == cudaErrorMemoryAllocation ? FSG_E_NOMEM : FSG_E_CUDA;                          \
        }                                                                                               \
    } while (0)

extern "C" int fsg_write_frame_async(fsg_ctx *c, const char *filename, int use_binary)
{
    if (!c || !filename) return FSG_E_INVALID;
    if (c->steps < 1) { c->err = "fsg_write_frame_async: no step taken since upload"; return FSG_E_STATE; }
    const int64_t n = c->n;
    if (n > 0x7fffffff / 3) { c->err = "fsg_write_frame_async: too many particles for the legacy VTK point-cloud writer"; return FSG_E_INVALID; }
    CUF(c, cudaSetDevice(c->device));
    FsgFrameWriter *w = c->frame_writer;
    if (!w) {
        w = new (std::nothrow) FsgFrameWriter();
        if (!w) return FSG_E_NOMEM;
        c->frame_writer = w;
        w->device = c->device;
        CUF(c, cudaStreamCreateWithFlags(&w->copy, cudaStreamNonBlocking));
        for (int k = 0; k < 2; k++) {
            CUF(c, cudaEventCreateWithFlags(&w->exported[k], cudaEventDisableTiming));
            CUF(c, cudaEventCreateWithFlags(&w->copied[k], cudaEventDisableTiming | cudaEventBlockingSync));
        }
        w->th = std::thread([w] { w->run(); });
    }
    int slot;
    {
        std::unique_lock<std::mutex> lk(w->mu);
        w->cv.wait(lk, [&] { return !w->slot_busy[0] || !w->slot_busy[1]; });       // back-pressure: at most two frames in flight
        slot = w->slot_busy[0] ? 1 : 0;
        w->slot_busy[slot] = true;
    }
    auto release = [&] { std::lock_guard<std::mutex> lk(w->mu); w->slot_busy[slot] = false; };
    if (w->cap[slot] < n) {
        if (w->dev[slot]) cudaFree(w->dev[slot]);
        if (w->host[slot]) cudaFreeHost(w->host[slot]);
        w->dev[slot] = nullptr; w->host[slot] = nullptr; w->cap[slot] = 0;
        const size_t bytes = (size_t)n * 20 + 256;
        if (cudaMalloc(&w->dev[slot], bytes) != cudaSuccess || cudaHostAlloc((void **)&w->host[slot], bytes, cudaHostAllocDefault) != cudaSuccess) {
            cudaGetLastError();
            release();
            c->err = "fsg_write_frame_async: cannot allocate the staging buffers";
            return FSG_E_NOMEM;
        }
        w->cap[slot] = n;
    }
    float *ds = w->dev[slot], *da = ds + 3 * n, *db = da + n;
    cudaError_t e = fsg_launch_export_viz(n, c->A.posd, c->keysA, ds, da, db, c->stream);       // mykernel2's export, FluidGPU.cu:410-414
    c->launches++;
    if (e == cudaSuccess && c->cfg.model == FSG_MODEL_UNIDYN) {                                   // a3 = mass, b3 = |diffusion|^2 (FluidGPU-unidyn.cu:465-466)
        e = c->cfg.unidyn_adapt ? fsg_launch_export_mass(n, c->A.mix, da, c->stream) : fsg_launch_fill((int *)da, 0x3f800000, n, c->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(db, c->vizb, sizeof(float) * (size_t)n, cudaMemcpyDeviceToDevice, c->stream);
        c->launches++;
    }
    if (e == cudaSuccess) e = cudaEventRecord(w->exported[slot], c->stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(w->copy, w->exported[slot], 0);
    if (e == cudaSuccess) e = cudaMemcpyAsync(w->host[slot], ds, (size_t)n * 20, cudaMemcpyDeviceToHost, w->copy);
    if (e == cudaSuccess) e = cudaEventRecord(w->copied[slot], w->copy);
    if (e != cudaSuccess) { release(); c->err = std::string("fsg_write_frame_async: ") + cudaGetErrorString(e); return FSG_E_CUDA; }
    {
        std::lock_guard<std::mutex> lk(w->mu);
        w->q.push_back(FsgFrameWriter::Job{filename, use_binary, n, slot, c->cfg.model == FSG_MODEL_UNIDYN});
    }
    w->cv.notify_all();
    return FSG_OK;
}

// Waits until every frame handed to fsg_write_frame_async is on disk; returns (and clears) the first error.
extern "C" int fsg_frame_wait(fsg_ctx *c, int64_t *frames_written)
{
    if (!c) return FSG_E_INVALID;
    FsgFrameWriter *w = c->frame_writer;
    if (!w) { if (frames_written) *frames_written = 0; return FSG_OK; }
    std::unique_lock<std::mutex> lk(w->mu);
    w->cv.wait(lk, [&] { return w->q.empty() && !w->busy; });
    if (frames_written) *frames_written = w->frames_written;
    const int rc = w->first_error;
    if (rc != FSG_OK) c->err = w->err;
    w->first_error = FSG_OK;
    return rc;
}
