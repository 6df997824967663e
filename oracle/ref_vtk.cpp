/*
 * ref_vtk.cpp — golden generator for the frame writer: calls the reference's UNMODIFIED write_point_mesh
 * (visit_writer.cpp:673, built into oracle/_ref/visit_writer.o) on arrays read from a section file.
 * TEST INFRASTRUCTURE ONLY; contains no reference source.
 *   ref_vtk <in.bin> <out.vtk> <binary 0|1>
 * Section file: "pts" f32[3n], then any number of f32 sections: scalars [n] or vectors [3n], in file order.
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "visit_writer.h"

int main(int argc, char **argv)
{
    if (argc != 4) { fprintf(stderr, "usage: %s in.bin out.vtk binary\n", argv[0]); return 2; }
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 2;
    std::vector<std::string> names;
    std::vector<std::vector<float>> data;
    for (;;) {
        char nm[17] = {0};
        int dtype;
        long long count;
        if (fread(nm, 1, 16, f) != 16) break;
        if (fread(&dtype, 4, 1, f) != 1 || fread(&count, 8, 1, f) != 1 || dtype != 0) return 2;
        std::vector<float> v((size_t)count);
        if (fread(v.data(), 4, (size_t)count, f) != (size_t)count) return 2;
        names.push_back(nm);
        data.push_back(v);
    }
    fclose(f);
    if (names.empty() || names[0] != "pts") return 2;
    int npts = (int)(data[0].size() / 3), nvars = (int)names.size() - 1;
    std::vector<int> dims;
    std::vector<const char *> vn;
    std::vector<float *> vp;
    for (int i = 1; i <= nvars; i++) { dims.push_back((int)(data[i].size() / npts)); vn.push_back(names[i].c_str()); vp.push_back(data[i].data()); }
    write_point_mesh(argv[2], atoi(argv[3]), npts, data[0].data(), nvars, dims.data(), vn.data(), vp.data());
    return 0;
}
