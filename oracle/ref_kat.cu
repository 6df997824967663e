/*
 * ref_kat.cu — known-answer generator: calls the reference's own host-compiled kernel(),
 * kernel_test(), kernel_derivative() and Particle::set_dens() (all __host__ __device__,
 * FluidGPU.cuh:43-47,165) from the UNMODIFIED object oracle/_ref/FluidGPU.o and prints them as
 * JSON with exact hex floats.  Runs on the CPU (no GPU needed).  TEST INFRASTRUCTURE ONLY.
 * Output is committed as tests/golden/kat_base.json by tests/golden/make_kat.py.
 */
#include <cstdio>
#include <cstdlib>
#include "FluidGPU.cuh"

int main(int argc, char **argv)
{
    int nr = argc > 1 ? atoi(argv[1]) : 257;
    printf("{\"cutoff\": %.17g, \"kernel\": [", (double)cutoff);
    for (int i = 0; i < nr; i++) {
        float r = (float)(0.13 * i / (nr - 1));
        printf("%s[\"%a\", \"%a\", \"%a\", \"%a\"]", i ? ", " : "", r, kernel(r), kernel_test(r), kernel_derivative(r));
    }
    printf("], \"set_dens\": [");
    const float xs[] = {0.f, 100.f, 1234.5f, 9550.f, 12000.f, 20000.f, 31415.9f};
    for (int b = 0; b < 2; b++)
        for (unsigned i = 0; i < sizeof(xs) / sizeof(xs[0]); i++) {
            Particle p(0.f, 0.f, 0.f, b != 0);
            p.set_dens(xs[i]);
            printf("%s[%d, \"%a\", \"%a\"]", (b || i) ? ", " : "", b, xs[i], p.dens);
        }
    printf("], \"sizeof_particle\": %d}\n", (int)sizeof(Particle));
    return 0;
}
