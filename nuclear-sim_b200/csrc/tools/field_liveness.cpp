// Developer tool (host build of the product's own single-source physics, g++): which PlantState fields are
// LIVE ON ENTRY to a step, i.e. read before they are overwritten?  Only those need to travel from HBM into the SM
// ahead of use, so the step kernel's software prefetch (hd.h nps_prefetch_*) is generated from this list.
// The answer is a performance hint: a wrong bit costs a cache miss, never a wrong result.
//
// Method: for every base state b and field f, perturb f three ways, run one plant_step, and compare every field
// against the unperturbed run bit for bit (f itself is exempt when the step leaves it untouched); if nothing differs for any perturbation, f was
// dead on entry in that trajectory.  Base states = the given snapshot advanced 0, 1, 7 and 40 steps.
//
//   g++ -O2 -std=c++17 -ffp-contract=off -I../plant field_liveness.cpp -o field_liveness
//   ./field_liveness snapshot.bin [...]        (snapshot.bin = n_state doubles then n_params doubles)
// prints one line of n_state characters: '1' live on entry, '0' dead on entry.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>
#include "plant_step.h"
using namespace nps;
constexpr int NS = sizeof(PlantState) / sizeof(double);
constexpr int NP = sizeof(PlantParams) / sizeof(double);

static StepInput input_for(int t) {
    StepInput in;
    static const int acts[] = {8, 0, 1, 8, 2, 3, 8, 4, 5, 6, 7, 8, 9, 10};
    in.action = acts[t % 14];
    in.magnitude = 0.5;
    in.z_heat = 0.3; in.z_ph = -0.2; in.u_ph[0] = 0.9; in.u_ph[1] = 0.8; in.u_ph[2] = 0.7;
    in.power_setpoint = NAN;
    return in;
}

int main(int argc, char** argv) {
    std::vector<char> live(NS, 0);
    for (int a = 1; a < argc; ++a) {
        FILE* fh = fopen(argv[a], "rb");
        if (!fh) { perror(argv[a]); return 1; }
        std::vector<double> buf(NS + NP);
        if (fread(buf.data(), sizeof(double), NS + NP, fh) != (size_t)(NS + NP)) { fprintf(stderr, "short read\n"); return 1; }
        fclose(fh);
        PlantParams prm; std::memcpy(&prm, buf.data() + NS, sizeof(prm));
        PlantState base; std::memcpy(&base, buf.data(), sizeof(base));
        int t = 0;
        const int stops[] = {0, 1, 7, 40};
        for (int s = 0; s < 4; ++s) {
            for (; t < stops[s]; ++t) plant_step(base, prm, input_for(t));
            PlantState ref = base;
            plant_step(ref, prm, input_for(t));
            const double* r = reinterpret_cast<const double*>(&ref);
            for (int f = 0; f < NS; ++f) {
                if (live[f]) continue;
                for (int pert = 0; pert < 3 && !live[f]; ++pert) {
                    PlantState st = base;
                    double* v = reinterpret_cast<double*>(&st);
                    const double x = v[f];
                    v[f] = (pert == 0) ? x * 1.37 + 0.61 : (pert == 1 ? (x == 0.0 ? 1.0 : 0.0) : NAN);
                    const double xin = v[f];
                    plant_step(st, prm, input_for(t));
                    // a field the step never touches keeps the perturbed value: that alone is not a read
                    if (std::memcmp(&v[f], &xin, sizeof(double)) == 0 && std::memcmp(&r[f], &x, sizeof(double)) == 0) v[f] = r[f];
                    if (std::memcmp(v, r, sizeof(double) * NS) != 0) live[f] = 1;
                }
            }
        }
    }
    for (int f = 0; f < NS; ++f) putchar(live[f] ? '1' : '0');
    putchar('\n');
    return 0;
}
