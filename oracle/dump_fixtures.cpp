// TEST INFRASTRUCTURE ONLY -- fixture extraction.
//
// Compiles the reference's own test translation unit (its fixture builders make_charge_square,
// make_waterbox, make_methanolbox; platforms/reference/tests/TestReferenceMPIDForce.cpp:64-880)
// with its main() renamed, runs the builders, and writes every parameter they set as JSON.
// The numbers therefore come from the reference's code, not from a transcription.
#define main reference_test_main
#include REF_TEST_FILE
#undef main
#include <cstdio>

static void dumpForce(FILE* f, const char* name, MPIDForce& force, const std::vector<Vec3>& pos, double box) {
    int n = force.getNumMultipoles();
    fprintf(f, "\"%s\": {\n \"n\": %d, \"box\": %.17g,\n \"positions\": [", name, n, box);
    for (int i = 0; i < n; i++) fprintf(f, "%s[%.17g, %.17g, %.17g]", i ? ", " : "", pos[i][0], pos[i][1], pos[i][2]);
    fprintf(f, "],\n \"multipoles\": [\n");
    for (int i = 0; i < n; i++) {
        double c, thole; std::vector<double> d, q, o, a; int ax, az, axx, ay;
        force.getMultipoleParameters(i, c, d, q, o, ax, az, axx, ay, thole, a);
        fprintf(f, "  {\"charge\": %.17g, \"dipole\": [%.17g, %.17g, %.17g], \"quadrupole\": [", c, d[0], d[1], d[2]);
        for (int k = 0; k < 6; k++) fprintf(f, "%s%.17g", k ? ", " : "", q[k]);
        fprintf(f, "], \"octopole\": [");
        for (int k = 0; k < 10; k++) fprintf(f, "%s%.17g", k ? ", " : "", o[k]);
        fprintf(f, "], \"axisType\": %d, \"atomZ\": %d, \"atomX\": %d, \"atomY\": %d, \"thole\": %.17g, \"alpha\": [%.17g, %.17g, %.17g], \"covalent\": [",
                ax, az, axx, ay, thole, a[0], a[1], a[2]);
        for (int t = 0; t < MPIDForce::CovalentEnd; t++) {
            std::vector<int> lst;
            force.getCovalentMap(i, (MPIDForce::CovalentType) t, lst);
            fprintf(f, "%s[", t ? ", " : "");
            for (size_t k = 0; k < lst.size(); k++) fprintf(f, "%s%d", k ? ", " : "", lst[k]);
            fprintf(f, "]");
        }
        fprintf(f, "]}%s\n", i+1 < n ? "," : "");
    }
    fprintf(f, " ]\n}");
}

int main(int argc, char** argv) {
    FILE* f = fopen(argc > 1 ? argv[1] : "fixtures.json", "w");
    fprintf(f, "{\n");
    {
        System system; MPIDForce* force = new MPIDForce(); std::vector<Vec3> pos;
        make_charge_square(2.0, pos, force, system);
        dumpForce(f, "charge_square", *force, pos, 2.0); system.addForce(force);
    }
    fprintf(f, ",\n");
    {
        System system; MPIDForce* force = new MPIDForce(); std::vector<Vec3> pos;
        make_waterbox(6, 2.0, force, pos, system);
        dumpForce(f, "water_dimer", *force, pos, 2.0); system.addForce(force);
    }
    fprintf(f, ",\n");
    {
        System system; MPIDForce* force = new MPIDForce(); std::vector<Vec3> pos;
        make_waterbox(375, 1.8643, force, pos, system);
        dumpForce(f, "water_375", *force, pos, 1.8643); system.addForce(force);
    }
    fprintf(f, ",\n");
    {
        System system; MPIDForce* force = new MPIDForce(); std::vector<Vec3> pos;
        double box = 24.61817*OpenMM::NmPerAngstrom;
        make_methanolbox(12, box, force, pos, system);
        dumpForce(f, "methanol_dimer", *force, pos, box); system.addForce(force);
    }
    fprintf(f, "\n}\n");
    fclose(f);
    return 0;
}
