// Drives the C++ mirror of PHDNavigator (monorfs_b200/host/phd_navigator.hpp) the way the reference's tests do
// (mono-rfs-lib/Test/PHDNavigatorTest.cs:85-126, Test/SimulationTest.cs:225-270), on the GPU through librbphd.so.
#include <cmath>
#include <cstdio>
#include <set>

#include "../../monorfs_b200/host/phd_navigator.hpp"

using namespace monorfs;

#define REQUIRE(cond)                                                           \
    do {                                                                        \
        if (!(cond)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); return 1; } \
    } while (0)

int main()
{
    rbphd_config cfg = DefaultPRM3DConfig();
    Pose3D origin = {{0, 0, 0, 1, 0, 0, 0}};
    PHDNavigator nav(cfg, origin, 5, false);
    REQUIRE(nav.Particles() == 5);

    // PredictInitial: empty map + one measurement -> exactly one birth with BirthCovariance / BirthWeight
    std::vector<PixelRangeMeasurement> z = {{10.0, -20.0, 1.5}};
    Map predicted = nav.PredictConditional(z, origin, Map());
    REQUIRE(predicted.size() == 1);
    REQUIRE(std::fabs(predicted[0].Weight - 0.05) < 1e-12);
    REQUIRE(std::fabs(predicted[0].Covariance[0] - 1e-2) < 1e-12 && predicted[0].Covariance[1] == 0.0);
    double alpha = 1.5 / std::sqrt(575.8156 * 575.8156 + 100.0 + 400.0);
    REQUIRE(std::fabs(predicted[0].Mean[0] - alpha * 10.0) < 1e-12);
    REQUIRE(std::fabs(predicted[0].Mean[2] - alpha * 575.8156) < 1e-12);

    // PredictKnown: an explored location gives no birth
    Map known(1);
    known[0].Weight = 1.0;
    for (int a = 0; a < 3; a++) known[0].Mean[a] = predicted[0].Mean[a];
    for (int a = 0; a < 9; a++) known[0].Covariance[a] = (a % 4 == 0) ? 1.0 : 0.0;
    REQUIRE(nav.PredictConditional(z, origin, known).size() == 1);

    // Correct + Prune on that map: miss-detected copy and the detection merge or stay apart, weights positive
    Map corrected = nav.CorrectConditional(z, origin, known);
    REQUIRE(corrected.size() == 2);
    Map pruned = nav.PruneModel(corrected);
    REQUIRE(!pruned.empty() && pruned.size() <= 2);

    // SimulationTest.resample: weights {.11,.28,.31,.01,.29}; particles 1, 2, 4 always survive, best is 2
    int seen0 = 0, seen3 = 0;
    const int iterations = 200;
    for (int it = 0; it < iterations; it++) {
        nav.CollapseParticles(5);
        for (int i = 0; i < 5; i++) { Pose3D p = {{(double)i, 0, 0, 1, 0, 0, 0}}; nav.SetVehiclePose(i, p); }
        nav.SetVehicleWeights({0.11, 0.28, 0.31, 0.01, 0.29});
        nav.ResampleParticles();
        std::vector<Pose3D> poses = nav.VehiclePoses();
        std::set<int> alive;
        for (const Pose3D& p : poses) alive.insert((int)p.State[0]);
        REQUIRE((int)poses[nav.BestParticle].State[0] == 2);
        REQUIRE(alive.count(1) && alive.count(2) && alive.count(4));
        seen0 += (int)alive.count(0);
        seen3 += (int)alive.count(3);
    }
    REQUIRE(seen0 < iterations && seen3 < iterations);

    // a few full frames: Update + SlamUpdate keep the weights normalised
    nav.CollapseParticles(5);
    const double reading[6] = {0, 0, 0.01, 0, 0.002, 0};
    for (int f = 0; f < 3; f++) {
        nav.Update(1.0 / 30.0, reading);
        nav.SlamUpdate({{10.0, -20.0, 1.5}, {-50.0, 30.0, 1.2}});
        double sum = 0;
        for (double w : nav.VehicleWeights()) sum += w;
        REQUIRE(std::fabs(sum - 1.0) < 1e-9);
        REQUIRE(nav.BestParticle >= 0 && nav.BestParticle < 5);
        REQUIRE(nav.BestMapModel().size() >= 1);
    }
    std::printf("host mirror ok\n");
    return 0;
}
