// ref_tps_kat.cpp -- TEST INFRASTRUCTURE.  Runs the reference's own beam-model code -- mqi::pbs<float> (beam data file,
// spot -> beamlet, histories per spot), treatment_machine_ion::create_beamsource / create_coordinate_transform and
// beam_module_ion -- UNMODIFIED from /root/reference on a plan built in memory, and prints what it produced.  The
// GDCM-backed mqi::dataset those headers read the RTPLAN through is replaced by oracle/ref_dataset_stub.hpp (GDCM is an
// absent third-party dependency); everything above it is the reference's code.
//
//   ref_tps_kat <machine file> <spots file: "E x y meterset" per line> <particles per history> <snout position>
//               <collimator> <gantry> <couch> <iso x> <iso y> <iso z>
//
// Output (stdout): "angles a0 a1 a2 a3", "trans x y z", then per beamlet
//   "spot <i> <histories> <E> <sigma E> <mean x6> <sigma x6>"
// the standard headers first (include guards keep them out of reach of the define below)
#include <algorithm>
#include <array>
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <functional>
#include <iostream>
#include <limits>
#include <map>
#include <numeric>
#include <queue>
#include <random>
#include <regex>
#include <sstream>
#include <string>
#include <tuple>
#include <valarray>
#include <vector>
#define protected public   // beamlet / pdf_Md keep their parameters protected; the harness only reads them
#include <moqui/base/mqi_treatment_machine_pbs.hpp>
#undef protected

typedef float R;

int
main(int argc, char* argv[]) {
    if (argc < 11) return 2;
    const std::string machine_file = argv[1], spots_file = argv[2];
    const float       pph = std::atof(argv[3]), snout = std::atof(argv[4]);
    struct spot_in { std::string e, x, y, w; };
    std::vector<spot_in> spots;
    {
        std::ifstream f(spots_file);
        std::string   line;
        while (std::getline(f, line)) {
            std::stringstream ss(line);
            spot_in           s;
            if (ss >> s.e >> s.x >> s.y >> s.w) spots.push_back(s);
        }
    }
    // IonBeamSequence item: one pair of ion control points per energy layer (the second of each pair carries zero
    // weights and is dropped by beam_module_ion), geometry on the first control point
    mqi::dataset              beam;
    std::vector<mqi::dataset*> cps;
    beam.set("ScanMode", { "MODULATED" });
    size_t i = 0;
    while (i < spots.size()) {
        size_t j = i;
        while (j < spots.size() && spots[j].e == spots[i].e) ++j;
        mqi::dataset* cp = new mqi::dataset;
        std::vector<std::string> xy, w, zero;
        for (size_t k = i; k < j; ++k) {
            xy.push_back(spots[k].x);
            xy.push_back(spots[k].y);
            w.push_back(spots[k].w);
            zero.push_back("0");
        }
        cp->set("ScanSpotTuneID", { "4.0" }).set("NominalBeamEnergy", { spots[i].e })
          .set("NumberOfScanSpotPositions", { std::to_string(j - i) }).set("ScanningSpotSize", { "10", "10" })
          .set("ScanSpotPositionMap", xy).set("ScanSpotMetersetWeights", w);
        if (cps.empty())
            cp->set("BeamLimitingDeviceAngle", { argv[5] }).set("GantryAngle", { argv[6] }).set("PatientSupportAngle", { argv[7] })
              .set("IsocenterPosition", { argv[8], argv[9], argv[10] }).set("SnoutPosition", { argv[4] });
        mqi::dataset* cp2 = new mqi::dataset(*cp);
        cp2->set("ScanSpotMetersetWeights", zero);
        cps.push_back(cp);
        cps.push_back(cp2);
        i = j;
    }
    for (auto* c : cps) beam.add("ctrl", c);
    // optional beamline devices: --rs-id ID... | --rs-wet WET DIST (RangeShifterSettingsSequence of the first control
    // point) and --block THICKNESS TRAY_DISTANCE x y x y ... (one IonBlockSequence item per --block)
    int n_rs = 0, n_blk = 0;
    for (int a = 11; a < argc;) {
        const std::string o = argv[a++];
        if (o == "--rs-id") {
            while (a < argc && argv[a][0] != '-') {
                mqi::dataset* r = new mqi::dataset;
                r->set("RangeShifterID", { argv[a++] });
                beam.add("rs", r);
                ++n_rs;
            }
        } else if (o == "--rs-wet" && a + 1 < argc) {
            mqi::dataset* ss = new mqi::dataset;
            ss->set("RangeShifterWaterEquivalentThickness", { argv[a] }).set("IsocenterToRangeShifterDistance", { argv[a + 1] });
            a += 2;
            cps[0]->add("rsss", ss);
            mqi::dataset* r = new mqi::dataset;
            r->set("RangeShifterID", { "unlisted" });
            beam.add("rs", r);
            ++n_rs;
        } else if (o == "--block" && a + 1 < argc) {
            mqi::dataset* b = new mqi::dataset;
            b->set("BlockThickness", { argv[a] }).set("IsocenterToBlockTrayDistance", { argv[a + 1] });
            a += 2;
            std::vector<std::string> xy;
            while (a < argc && std::string(argv[a]).compare(0, 2, "--") != 0) xy.push_back(argv[a++]);
            b->set("BlockNumberOfPoints", { std::to_string(xy.size() / 2) }).set("BlockData", xy);
            beam.add("blk", b);
            ++n_blk;
        }
    }
    beam.set("NumberOfRangeShifters", { std::to_string(n_rs) }).set("NumberOfBlocks", { std::to_string(n_blk) });

    mqi::pbs<R>                  machine(machine_file);
    mqi::coordinate_transform<R> pc = machine.create_coordinate_transform(&beam, mqi::IONPLAN);
    printf("angles %.9g %.9g %.9g %.9g\n", pc.angles[0], pc.angles[1], pc.angles[2], pc.angles[3]);
    printf("trans %.9g %.9g %.9g\n", pc.translation.x, pc.translation.y, pc.translation.z);
    // tps_env generates the histories 50 mm upstream of the snout (mqi_tps_env.hpp:975)
    mqi::beamsource<R> bs = machine.create_beamsource(&beam, mqi::IONPLAN, pc, pph, snout + 50.0f);
    for (size_t b = 0; b < bs.total_beamlets(); ++b) {
        const auto&            t  = bs[b];
        const mqi::beamlet<R>& bl = std::get<0>(t);
        printf("spot %zu %zu %.9g %.9g", b, (size_t) std::get<1>(t), bl.energy->mean_[0], bl.energy->sigma_[0]);
        for (int k = 0; k < 6; ++k) printf(" %.9g", bl.fluence->mean_[k]);
        for (int k = 0; k < 6; ++k) printf(" %.9g", bl.fluence->sigma_[k]);
        printf("\n");
    }
    // create_beamline: characterize_rangeshifter / characterize_aperture, sorted upstream first (tmi:238-276)
    mqi::beamline<R> line = machine.create_beamline(&beam, mqi::IONPLAN);
    for (mqi::geometry* g : line.get_geometries()) {
        if (g->geotype == mqi::RANGESHIFTER) {
            const mqi::rangeshifter* r = dynamic_cast<const mqi::rangeshifter*>(g);
            printf("geo rangeshifter %.9g %.9g %.9g %.9g %.9g %.9g\n", r->volume.x, r->volume.y, r->volume.z, g->pos.x, g->pos.y, g->pos.z);
        } else if (g->geotype == mqi::BLOCK) {
            const mqi::aperture* ap = dynamic_cast<const mqi::aperture*>(g);
            printf("geo block %.9g %.9g %.9g %.9g %.9g %.9g\n", ap->volume.x, ap->volume.y, ap->volume.z, g->pos.x, g->pos.y, g->pos.z);
        }
    }
    return 0;
}
