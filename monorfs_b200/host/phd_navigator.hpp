// phd_navigator.hpp -- host-side mirror of MonoRFS's PHDNavigator<PRM3DMeasurer, Pose3D, PixelRangeMeasurement>
// on top of the librbphd C ABI (include/rbphd.h).
//
// The reference's host language is C# (no toolchain in this image), so the class that a maintainer would
// write as `GpuPHDNavigator : Navigator<...>` is mirrored here in C++ with the same
// member names, argument meaning and error behaviour (mono-rfs-lib/SLAM/Navigators/PHDNavigator.cs:
// ctor :192, Update :295, SlamUpdate :323, WeightAlpha :373, ResampleParticles :724, ParticleDepleted :768,
// PredictConditional :793, CorrectConditional :829, PruneModel :913, CollapseParticles :233,
// ResetMapModel :271).  Failures surface as std::runtime_error carrying the library's message -- the
// counterpart of the InvalidOperationException with Data["module"] that Simulation.Update catches
// (ISAM2Navigator.cs:239-247, Simulation.cs:655-670).  Trajectory / map history (WayPoints, WayMaps)
// stays on the host side, as in Navigator.cs:258-272.
#pragma once
#include <cmath>
#include <cstring>
#include <functional>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/rbphd.h"

namespace monorfs {

struct Gaussian {           // Gaussian.cs:40-59
    double Weight;
    double Mean[3];
    double Covariance[9];   // row-major 3x3
};
typedef std::vector<Gaussian> Map;                       // Map.cs:41 (enumeration = insertion order)
struct PixelRangeMeasurement { double X, Y, Range; };    // PixelRangeMeasurement.cs:96-99
struct Pose3D { double State[7]; };                      // Pose3D.cs:110-113 (x, y, z, qw, qx, qy, qz)

class PHDNavigator {
public:
    int ParticleCount;
    bool OnlyMapping;
    int BestParticle = 0;
    bool PerfectStill = false;          // Config.PerfectStill
    std::vector<std::vector<Pose3D> > WayTrajectories;   // host-side history (Navigator.cs:258-266)

    // draws the 6 N(0,1) values per particle of TrackVehicle.UpdateNoisy (TRK:95-97); the C# host plugs in
    // Util.Gaussian.Next() here so the random stream stays the reference's
    std::function<double()> GaussianDraw;
    std::function<double()> UniformDraw;   // Util.Uniform.Next() of PHD:727 (float-valued)

    PHDNavigator(const rbphd_config& config, const Pose3D& initial, int particlecount, bool onlymapping = false,
                 const rbphd_limits* limits = nullptr)
        : ParticleCount(particlecount), OnlyMapping(onlymapping), cfg_(config), rng_(20261018)
    {
        rbphd_limits lim;
        std::memset(&lim, 0, sizeof lim);
        if (limits) lim = *limits;
        if (lim.max_particles < particlecount) lim.max_particles = particlecount;
        nav_ = rbphd_new(&cfg_, &lim);
        if (!nav_) throw std::runtime_error(std::string("librbphd: ") + rbphd_last_error(nullptr));
        GaussianDraw = [this]() { return (double)(float)normal_(rng_); };
        UniformDraw = [this]() { return (double)(float)uniform_(rng_); };
        ref_pose_ = initial;
        reset(initial, Map(), onlymapping ? 1 : particlecount);   // PHD:201-207
    }
    ~PHDNavigator() { rbphd_delete(nav_); }
    PHDNavigator(const PHDNavigator&) = delete;
    PHDNavigator& operator=(const PHDNavigator&) = delete;

    int Particles() const { return rbphd_particle_count(nav_); }
    void SetReferencePose(const Pose3D& p) { ref_pose_ = p; }   // RefVehicle.Pose

    // PHD:233-236
    void CollapseParticles(int particlecount) { reset(ref_pose_, BestMapModel(), particlecount); }
    void StartSlam() { OnlyMapping = false; CollapseParticles(ParticleCount); }      // PHD:214-217
    void StartMapping() { OnlyMapping = true; CollapseParticles(1); }                // PHD:224-227
    void ResetMapModel() { check(rbphd_clear_maps(nav_)); }                          // PHD:271-276

    // PHD:295-314
    void Update(double elapsedSeconds, const double reading[6])
    {
        if (OnlyMapping) {
            check(rbphd_set_pose(nav_, 0, ref_pose_.State));
        }
        else {
            const int P = Particles();
            gauss_.resize(6 * (size_t)P);
            for (double& g : gauss_) g = GaussianDraw();
            check(rbphd_update(nav_, reading, elapsedSeconds, gauss_.data(), PerfectStill ? 1 : 0));
        }
    }

    // PHD:323-362
    void SlamUpdate(const std::vector<PixelRangeMeasurement>& measurements)
    {
        z_.resize(3 * measurements.size() + 3);
        for (size_t k = 0; k < measurements.size(); k++) {
            z_[3 * k] = measurements[k].X; z_[3 * k + 1] = measurements[k].Y; z_[3 * k + 2] = measurements[k].Range;
        }
        // two steps, so that the uniform of the wheel is drawn exactly when the reference draws it (only when
        // ParticleDepleted(), PHD:355-357 -> PHD:727)
        int best = 0, depleted = 0;
        check(rbphd_slam_update_begin(nav_, z_.data(), (int)measurements.size(), OnlyMapping ? 1 : 0, &best, &depleted));
        if (depleted) check(rbphd_slam_update_finish(nav_, UniformDraw(), &best));
        BestParticle = best;
        LastResampled = depleted != 0;
    }
    bool LastResampled = false;

    void ResampleParticles()   // PHD:724-760
    {
        check(rbphd_resample(nav_, UniformDraw()));
        check(rbphd_get_best(nav_, &BestParticle));
    }
    bool ParticleDepleted()    // PHD:768-777
    {
        int d = 0;
        check(rbphd_particle_depleted(nav_, &d));
        return d != 0;
    }

    std::vector<double> VehicleWeights()
    {
        const double* w; int n;
        check(rbphd_get_weights(nav_, &w, &n));
        return std::vector<double>(w, w + n);
    }
    void SetVehicleWeights(const std::vector<double>& w) { check(rbphd_set_weights(nav_, w.data())); }
    std::vector<Pose3D> VehiclePoses()
    {
        const double* p; int n;
        check(rbphd_get_poses(nav_, &p, &n));
        std::vector<Pose3D> out(n);
        for (int i = 0; i < n; i++) std::memcpy(out[i].State, p + 7 * i, sizeof out[i].State);
        return out;
    }
    void SetVehiclePose(int i, const Pose3D& p) { check(rbphd_set_pose(nav_, i, p.State)); }
    Map MapModel(int i)
    {
        const double *w, *m, *P; int n;
        check(rbphd_get_map(nav_, i, &w, &m, &P, &n));
        return to_map(w, m, P, n);
    }
    void SetMapModel(int i, const Map& map)
    {
        flatten(map);
        check(rbphd_set_map(nav_, i, (int)map.size(), fw_.data(), fm_.data(), fP_.data()));
    }
    Map BestMapModel() { return Particles() > 0 ? MapModel(BestParticle) : Map(); }   // PHD:155-161
    std::vector<int> LastAncestors()
    {
        const int* a; int n;
        check(rbphd_get_ancestors(nav_, &a, &n));
        return std::vector<int>(a, a + n);
    }

    // the per-particle public methods (PHD:793, 829, 913, 373)
    Map PredictConditional(const std::vector<PixelRangeMeasurement>& measurements, const Pose3D& pose, const Map& model)
    {
        flatten(model); pack(measurements);
        const double *w, *m, *P; int n;
        check(rbphd_stage_predict(nav_, pose.State, (int)model.size(), fw_.data(), fm_.data(), fP_.data(), z_.data(),
                                  (int)measurements.size(), &w, &m, &P, &n));
        return to_map(w, m, P, n);
    }
    Map CorrectConditional(const std::vector<PixelRangeMeasurement>& measurements, const Pose3D& pose, const Map& model,
                           double gate_radius = NAN)
    {
        flatten(model); pack(measurements);
        if (std::isnan(gate_radius)) gate_radius = cfg_.density_distance_threshold;
        const double *w, *m, *P; int n;
        check(rbphd_stage_correct(nav_, pose.State, (int)model.size(), fw_.data(), fm_.data(), fP_.data(), z_.data(),
                                  (int)measurements.size(), gate_radius, &w, &m, &P, &n));
        return to_map(w, m, P, n);
    }
    Map PruneModel(const Map& model)
    {
        flatten(model);
        const double *w, *m, *P; int n;
        check(rbphd_stage_prune(nav_, (int)model.size(), fw_.data(), fm_.data(), fP_.data(), &w, &m, &P, &n));
        return to_map(w, m, P, n);
    }
    double WeightAlpha(const std::vector<PixelRangeMeasurement>& measurements, const Map& predicted,
                       const Map& corrected, const Pose3D& pose)
    {
        pack(measurements);
        flatten(predicted);
        std::vector<double> pw = fw_, pm = fm_, pP = fP_;
        flatten(corrected);
        double out[7];
        check(rbphd_stage_weight_alpha(nav_, pose.State, z_.data(), (int)measurements.size(), (int)predicted.size(),
                                       pw.data(), pm.data(), pP.data(), (int)corrected.size(), fw_.data(), fm_.data(),
                                       fP_.data(), out));
        return out[0];
    }
    // the static likelihood functions over a landmark list (PHD:395-406, 526-532)
    double SetLikelihood(const std::vector<PixelRangeMeasurement>& measurements, const Map& map, const Pose3D& pose)
    {
        pack(measurements);
        flatten(map);
        double v = 0;
        check(rbphd_set_likelihood(nav_, pose.State, (int)map.size(), fm_.data(), z_.data(), (int)measurements.size(), &v));
        return v;
    }
    double QuasiSetLogLikelihood(const std::vector<PixelRangeMeasurement>& measurements, const Map& map,
                                 const Pose3D& pose)
    {
        pack(measurements);
        flatten(map);
        double v = 0;
        check(rbphd_quasi_set_loglikelihood(nav_, pose.State, (int)map.size(), fm_.data(), z_.data(),
                                            (int)measurements.size(), &v));
        return v;
    }

private:
    void reset(const Pose3D& pose, const Map& model, int particlecount)   // PHD:245-266
    {
        flatten(model);
        check(rbphd_reset(nav_, particlecount, pose.State, (int)model.size(), fw_.data(), fm_.data(), fP_.data()));
        BestParticle = 0;
    }
    void check(int code)
    {
        if (code != RBPHD_OK)
            throw std::runtime_error("librbphd error " + std::to_string(code) + ": " + rbphd_last_error(nav_));
    }
    void flatten(const Map& map)
    {
        size_t n = map.size();
        fw_.assign(n + 1, 0); fm_.assign(3 * n + 3, 0); fP_.assign(9 * n + 9, 0);
        for (size_t i = 0; i < n; i++) {
            fw_[i] = map[i].Weight;
            std::memcpy(&fm_[3 * i], map[i].Mean, 3 * sizeof(double));
            std::memcpy(&fP_[9 * i], map[i].Covariance, 9 * sizeof(double));
        }
    }
    void pack(const std::vector<PixelRangeMeasurement>& ms)
    {
        z_.assign(3 * ms.size() + 3, 0);
        for (size_t k = 0; k < ms.size(); k++) { z_[3 * k] = ms[k].X; z_[3 * k + 1] = ms[k].Y; z_[3 * k + 2] = ms[k].Range; }
    }
    static Map to_map(const double* w, const double* m, const double* P, int n)
    {
        Map map(n);
        for (int i = 0; i < n; i++) {
            map[i].Weight = w[i];
            std::memcpy(map[i].Mean, m + 3 * i, 3 * sizeof(double));
            std::memcpy(map[i].Covariance, P + 9 * i, 9 * sizeof(double));
        }
        return map;
    }

    rbphd_config cfg_;
    rbphd_navigator* nav_ = nullptr;
    Pose3D ref_pose_;
    std::vector<double> gauss_, z_, fw_, fm_, fP_;
    std::mt19937_64 rng_;
    std::normal_distribution<double> normal_{0.0, 1.0};
    std::uniform_real_distribution<double> uniform_{0.0, 1.0};
};

// Config.SetPRM3DDefaults (Config.cs:238-263) + PRM3DMeasurer() (PRM3DMeasurer.cs:70-73)
inline rbphd_config DefaultPRM3DConfig()
{
    rbphd_config c;
    std::memset(&c, 0, sizeof c);
    c.model = 0; c.max_quantity = 600; c.gate_metric = 0; c.nthreads = 8;
    c.R[0] = 2.0; c.R[4] = 2.0; c.R[8] = 1e-3;
    for (int i = 0; i < 3; i++) { c.Q[i * 7] = 5e-3; c.Q[(i + 3) * 7] = 2e-4; }
    c.pd = 0.9; c.clutter = 3e-7;
    c.birth_cov[0] = c.birth_cov[4] = c.birth_cov[8] = 1e-2;
    c.birth_weight = 0.05; c.min_weight = 1e-3; c.merge_threshold = 0.3; c.exploration_threshold = 1e-5;
    c.density_distance_threshold = 0.5; c.min_effective_particle = 0.1;
    for (int i = 0; i < 3; i++) c.visibility_ramp[i] = 3 * std::sqrt(c.R[i * 4]);
    const double meas[7] = {575.8156, 0.1, 2.0, -320, -240, 640, 480};
    std::memcpy(c.measurer, meas, sizeof meas);
    return c;
}

}  // namespace monorfs
