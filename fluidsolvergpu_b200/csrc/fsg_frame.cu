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
#include <string.h>
#include <string>
#include <vector>

namespace {
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
            for (int64_t j = 0; j < (int64_t)npts * vardim[i]; j++) o.put_float(vars[i][j]);
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
            for (int64_t j = 0; j < (int64_t)npts * dim; j++) o.put_float(vars[i][j]);
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
    for (int64_t i = 0; i < 3ll * npts; i++) o.put_float(pts[i]);
    o.new_section();
    snprintf(s, sizeof s, "CELLS %d %d\n", npts, 2 * npts);
    o.str(s);
    for (int i = 0; i < npts; i++) { o.put_int(1); o.put_int(i); o.end_line(); }
    o.new_section();
    snprintf(s, sizeof s, "CELL_TYPES %d\n", npts);
    o.str(s);
    for (int i = 0; i < npts; i++) { o.put_int(1 /* VISIT_VERTEX */); o.end_line(); }
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
