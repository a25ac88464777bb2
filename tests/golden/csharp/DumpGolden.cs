// DumpGolden.cs -- pins the C++ oracle and the CUDA path against the REAL C# reference (afalchetti/monorfs).
//
// Neither Mono nor .NET exists in the image this repository is built in, so the oracle is "parity unpinned"
// against the C# for the items the reference's own NUnit tests do not cover (oracle/README.md D1-D9: Accord's
// KD-tree radius metric and enumeration order, List.Sort tie order, SVD pseudo-inverse thresholds, WeightAlpha).
// This harness closes that gap for anyone who has the reference built:
//
//   python tests/golden/csharp/export_inputs.py                      # writes inputs_slam_small.txt (seeded scene)
//   mcs -r:mono-rfs-lib.dll -r:Accord.Math.dll -r:AForge.Math.dll -r:MonoGame.Framework.dll \
//       tests/golden/csharp/DumpGolden.cs -out:DumpGolden.exe
//   mono DumpGolden.exe tests/golden/csharp/inputs_slam_small.txt tests/golden/csharp_slam_small.txt
//   python -m pytest tests/test_csharp_golden.py                      # oracle (CPU) and CUDA path (GPU) vs the C#
//
// What it runs, on the inputs of the file: for every frame the particle pose update (Pose3D.AddOdometry with
// the reading, then with dt * C g for the supplied N(0,1) draws g: TrackVehicle.UpdateNoisy, TRK:89-102, with
// the random draw replaced by the supplied one), then PHDNavigator.SlamUpdate (PHD:323-362) unchanged; and for
// particle 0 of frame 0 the per-stage public methods PredictConditional / CorrectConditional / PruneModel /
// WeightAlpha.  The wheel's uniform cannot be injected into AForge's generator, so the harness seeds Util.Uniform
// per frame, records what Next() returns and re-seeds: the value travels in the output and is fed to the oracle.
// Numbers are written with the round-trip format "r".
using System;
using System.Collections.Generic;
using System.Globalization;
using System.IO;
using System.Linq;
using Accord.Math;
using Accord.Math.Decompositions;
using Microsoft.Xna.Framework;
using monorfs;

using PHDNavigator     = monorfs.PHDNavigator<monorfs.PRM3DMeasurer, monorfs.Pose3D, monorfs.PixelRangeMeasurement>;
using SimulatedVehicle = monorfs.SimulatedVehicle<monorfs.PRM3DMeasurer, monorfs.Pose3D, monorfs.PixelRangeMeasurement>;

public static class DumpGolden
{
	static Dictionary<string, double[]> ReadInputs(string path)
	{
		var data = new Dictionary<string, double[]>();
		foreach (string line in File.ReadLines(path)) {
			string[] tok = line.Split(new char[] {' '}, StringSplitOptions.RemoveEmptyEntries);
			if (tok.Length < 2 || tok[0].StartsWith("#")) { continue; }
			int n = int.Parse(tok[1], CultureInfo.InvariantCulture);
			data[tok[0]] = tok.Skip(2).Take(n).Select(s => double.Parse(s, CultureInfo.InvariantCulture)).ToArray();
		}
		return data;
	}

	static void Put(TextWriter w, string name, IEnumerable<double> values)
	{
		double[] v = values.ToArray();
		w.WriteLine(name + " " + v.Length + " " + string.Join(" ", v.Select(x => x.ToString("r", CultureInfo.InvariantCulture))));
	}

	static double[][] Square(double[] flat, int n)
	{
		double[][] m = new double[n][];
		for (int i = 0; i < n; i++) { m[i] = flat.Skip(i * n).Take(n).ToArray(); }
		return m;
	}

	static Map ToMap(double[] w, double[] m, double[] P)
	{
		Map map = new Map(3);
		for (int i = 0; i < w.Length; i++) {
			map.Add(new Gaussian(m.Skip(3 * i).Take(3).ToArray(), Square(P.Skip(9 * i).Take(9).ToArray(), 3), w[i]));
		}
		return map;
	}

	static void PutMap(TextWriter w, string prefix, IMap map)
	{
		var comps = new List<Gaussian>();
		foreach (Gaussian g in map) { comps.Add(g); }    // enumeration order = the order parity is defined on
		Put(w, prefix + "_w", comps.Select(g => g.Weight));
		Put(w, prefix + "_m", comps.SelectMany(g => g.Mean));
		Put(w, prefix + "_P", comps.SelectMany(g => g.Covariance.SelectMany(r => r)));
	}

	// Util.RandomGaussianVector (UTIL:173-202) with the canonical vector supplied instead of drawn
	static double[] CorrelatedNoise(double[][] covariance, double[] canonical)
	{
		for (int i = 0; i < covariance.Length; i++) {
			if (covariance[i][i] < 1e-40) { covariance[i][i] = 1e-40; }
		}
		var       cholesky = new CholeskyDecomposition(covariance.ToMatrix());
		double[]  sqrtdiag = cholesky.Diagonal;
		for (int i = 0; i < sqrtdiag.Length; i++) { sqrtdiag[i] = Math.Sqrt(sqrtdiag[i]); }
		double[,] covroot = cholesky.LeftTriangularFactor.MultiplyByDiagonal(sqrtdiag);
		return new double[canonical.Length].Add(covroot.Multiply(canonical));
	}

	public static int Main(string[] args)
	{
		if (args.Length != 2) { Console.Error.WriteLine("usage: DumpGolden inputs.txt outputs.txt"); return 2; }
		var inp = ReadInputs(args[0]);
		int P = (int) inp["P"][0], M = (int) inp["M"][0], frames = (int) inp["frames"][0];
		double dt = inp["dt"][0];

		Config.SetPRM3DDefaults();
		Config.MeasurementCovariance = Square(inp["R"], 3);
		Config.MotionCovariance      = Square(inp["Q"], 6);
		Config.DetectionProbability  = inp["pd"][0];
		Config.ClutterDensity        = inp["clutter"][0];
		Config.NavigatorPD           = inp["pd"][0];
		Config.NavigatorClutterDensity = inp["clutter"][0];
		Config.VisibilityRamp        = inp["visibility_ramp"];
		Config.BirthCovariance       = Square(inp["birth_cov"], 3);
		Config.BirthWeight           = inp["birth_weight"][0];
		Config.MinWeight             = inp["min_weight"][0];
		Config.MaxQuantity           = (int) inp["max_quantity"][0];
		Config.MergeThreshold        = inp["merge_threshold"][0];
		Config.ExplorationThreshold  = inp["exploration_threshold"][0];
		Config.DensityDistanceThreshold = inp["density_distance_threshold"][0];
		Config.MinEffectiveParticle  = inp["min_effective_particle"][0];

		double[] ms = inp["measurer"];   // focal, range min, range max, film left, top, width, height
		var measurer = new PRM3DMeasurer(ms[0], new Rectangle((int) ms[3], (int) ms[4], (int) ms[5], (int) ms[6]),
		                                 new AForge.Range((float) ms[1], (float) ms[2]));
		var vehicle  = new SimulatedVehicle(new Pose3D(inp["poses0"].Take(7).ToArray()), new List<double[]>(), measurer);
		var nav      = new PHDNavigator(vehicle, P, false);

		Map map0 = ToMap(inp["map_w"], inp["map_m"], inp["map_P"]);
		for (int i = 0; i < P; i++) {
			nav.VehicleParticles[i].Pose = new Pose3D(inp["poses0"].Skip(7 * i).Take(7).ToArray());
			nav.MapModels[i]             = new Map(map0);
			nav.VehicleWeights[i]        = 1.0 / P;
		}

		using (var w = new StreamWriter(args[1])) {
			w.WriteLine("# written by DumpGolden.cs from " + Path.GetFileName(args[0]));
			// per-stage outputs of particle 0 on frame 0's measurements (prior map, initial pose)
			{
				var z0 = new List<PixelRangeMeasurement>();
				for (int k = 0; k < M; k++) { z0.Add(new PixelRangeMeasurement(inp["z0"][3 * k], inp["z0"][3 * k + 1], inp["z0"][3 * k + 2])); }
				var pose0     = nav.VehicleParticles[0];
				Map predicted = nav.PredictConditional(z0, pose0, new Map(map0), new List<double[]>());
				Map corrected = nav.CorrectConditional(z0, pose0, predicted);
				Map pruned    = nav.PruneModel(corrected);
				PutMap(w, "stage_predicted", predicted);
				PutMap(w, "stage_corrected", corrected);
				PutMap(w, "stage_pruned", pruned);
				Put(w, "stage_alpha", new double[] { nav.WeightAlpha(z0, predicted, pruned, pose0) });
				Put(w, "stage_setloglik", new double[] { PHDNavigator.SetLogLikelihood(z0, pruned.BestMapEstimate, pose0) });
			}

			TimeSpan total = TimeSpan.Zero, step = TimeSpan.FromSeconds(dt);
			for (int f = 0; f < frames; f++) {
				total += step;
				var time = new GameTime(total, step);
				double[] reading = inp["reading" + f], gauss = inp["gauss" + f], zf = inp["z" + f];
				for (int i = 0; i < P; i++) {   // TRK:89-102 with the draw supplied
					var v = nav.VehicleParticles[i];
					v.Pose = v.Pose.AddOdometry(reading);
					double[] noise = dt.Multiply(CorrelatedNoise(v.MotionCovariance, gauss.Skip(6 * i).Take(6).ToArray()));
					v.Pose = v.Pose.AddOdometry(noise);
				}
				var z = new List<PixelRangeMeasurement>();
				for (int k = 0; k < zf.Length / 3; k++) { z.Add(new PixelRangeMeasurement(zf[3 * k], zf[3 * k + 1], zf[3 * k + 2])); }

				Util.Uniform.SetSeed(1000 + f);
				double u = (double) Util.Uniform.Next();      // what PHD:727 will draw if it resamples
				Util.Uniform.SetSeed(1000 + f);
				double[] before = (double[]) nav.VehicleWeights.Clone();
				nav.SlamUpdate(time, z);

				bool resampled = nav.VehicleWeights.All(x => x == 1.0 / P) && !before.All(x => x == 1.0 / P);
				Put(w, "u" + f, new double[] { u });
				Put(w, "best" + f, new double[] { nav.BestParticle });
				Put(w, "res" + f, new double[] { resampled ? 1 : 0 });
				Put(w, "w" + f, nav.VehicleWeights);
				Put(w, "counts" + f, nav.MapModels.Select(m => (double) m.Count));
				Put(w, "poses" + f, nav.VehicleParticles.SelectMany(v => v.Pose.State));
			}
			for (int i = 0; i < P; i++) { PutMap(w, "final" + i, nav.MapModels[i]); }
		}
		return 0;
	}
}
