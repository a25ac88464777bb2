"""Record / replay wire format of MonoRFS: the files a `Simulation` saves into
`data.zip` (UI/Simulation.cs:391-488) and `RecordVehicle.FromFile` / `Viewer.FromFiles` / `postanalysis` read
back (SLAM/Vehicles/RecordVehicle.cs:244-347, Util/FileParser.cs:56-340), plus the scene (`.world`,
SLAM/Vehicles/Vehicle.cs:503-524, SimulatedVehicle.cs:346-385) and command (`.in`, FileParser.cs:263-275) inputs
of a simulation run.  Host-side text / zip handling only (the reference does this in C#); numbers are written
with the reference's "g6" format unless `lossless` is asked for (then repr(): every FileParser call is a plain
double.Parse, so longer digit strings read back fine in the C#).

Archive members:
  scene.world       pose / params / landmarks dictionary, children indented by one tab
  trajectory.out    one line per frame: time then the 7 pose state values (groundtruth)
  odometry.out      one line per frame: time then the 6 odometry values
  estimate.out      frames separated by "\\n|\\n": a time line, then the best particle's waypoints as trajectory lines
  maps.out          frames separated by "\\n|\\n": a time line, then one Gaussian per line  w;m0 m1 m2;c00 c01 .. c22
  vismaps.out       the same for the groundtruth visible map
  measurements.out  one line per frame: time:x y r;x y r;...
  tags.out          time message
  config.cfg        free text
"""
import io
import zipfile
from dataclasses import dataclass, field

import numpy as np

FRAME_SEP = "\n|\n"


def g6(x):
    """C# double.ToString("g6")."""
    x = float(x)
    if x != x:
        return "NaN"
    if x in (float("inf"), float("-inf")):
        return "Infinity" if x > 0 else "-Infinity"
    s = "%.6g" % x
    return "0" if s == "-0" else s


def _fmt(lossless):
    return (lambda v: repr(float(v))) if lossless else g6


def _doubles(text):
    return np.array([float(t) for t in text.split()], dtype=np.float64)


# ------------------------------------------------------------------ scene (.world) and commands (.in)
def scene_to_text(pose, measurer, landmarks, lossless=False):
    """Vehicle.ToString (Vehicle.cs:513-524): pose / params / landmarks."""
    f = _fmt(lossless)
    lm = np.asarray(landmarks, dtype=np.float64).reshape(-1, 3)
    return ("pose\n\t" + " ".join(f(v) for v in pose) + "\n" +
            "params\n\t" + " ".join(f(v) for v in measurer) + "\n" +
            "landmarks\n\t" + "\n\t".join(" ".join(f(v) for v in l) for l in lm) + "\n")


def parse_dictionary(text):
    """Util.ParseDictionary (Util.cs:232-264): keys at column 0, children behind one tab."""
    out, key = {}, None
    lines = text.replace("\r\n", "\n").replace("\r", "\n").split("\n")
    if lines and (not lines[0] or lines[0][0].isspace()):
        return out
    for line in lines:
        if not line.strip():
            continue
        if line[0] != "\t":
            key = line
            out[key] = []
        else:
            out[key].append(line[1:])
    return out


def parse_scene(text):
    """SimulatedVehicle.FromFile (SimulatedVehicle.cs:346-385): (pose[7], measurer[7] or None, landmarks[n,3])."""
    d = parse_dictionary(text)
    pose = _doubles(d["pose"][0])
    key = "focal" if "focal" in d else ("params" if "params" in d else None)
    measurer = _doubles(d[key][0]) if key else None
    lm = [_doubles(l) for l in d.get("landmarks", [])]
    for l in lm:
        if len(l) != 3:
            raise ValueError("Map landmarks must be 3D")
    return pose, measurer, np.array(lm, dtype=np.float64).reshape(-1, 3)


def commands_to_text(commands, lossless=False):
    """One command per line: the 6 odometry values and the SLAM / mapping switch (-1, 0, 1)."""
    f = _fmt(lossless)
    return "\n".join(" ".join(f(v) for v in c) for c in commands) + "\n"


def parse_commands(text):
    return [_doubles(line) for line in text.splitlines() if line.strip()]


# ------------------------------------------------------------------ histories
def timed_array_to_text(rows, lossless=False):
    """Simulation.SerializeWayPoints (Simulation.cs:225-231)."""
    f = _fmt(lossless)
    return "\n".join(g6(t) + " " + " ".join(f(v) for v in vals) for t, vals in rows)


def parse_timed_array(lines, dim):
    """FileParser.TimedArrayFromDescriptor (FileParser.cs:104-120)."""
    out = []
    for line in lines:
        if not line.strip():
            continue
        v = _doubles(line)
        if len(v) != dim + 1:
            raise ValueError("wrong state dimension")
        out.append((float(v[0]), v[1:]))
    return out


def gaussian_to_text(w, m, P, lossless=False):
    """Gaussian.ToString (Gaussian.cs:391-425): weight;mean;covariance row-major."""
    f = _fmt(lossless)
    return f(w) + ";" + " ".join(f(v) for v in np.ravel(m)) + ";" + " ".join(f(v) for v in np.ravel(P))


def parse_gaussian(text):
    """FileParser.ParseGaussianDescriptor (FileParser.cs:302-340)."""
    parts = text.split(";")
    w = float(parts[0])
    m = np.array([float(t) for t in parts[1].split(" ")])
    c = np.array([float(t) for t in parts[2].split(" ")])
    if len(c) != len(m) * len(m):
        raise ValueError("covariance has the wrong size")
    return w, m, c.reshape(len(m), len(m))


def map_history_to_text(history, lossless=False):
    """Simulation.SerializedMaps (Simulation.cs:199-207)."""
    frames = []
    for t, (w, m, P) in history:
        m, P = np.asarray(m).reshape(-1, 3), np.asarray(P).reshape(-1, 3, 3)
        frames.append(g6(t) + "\n" + "\n".join(gaussian_to_text(w[i], m[i], P[i], lossless) for i in range(len(w))))
    return FRAME_SEP.join(frames)


def parse_map_history(text, dim=3):
    """FileParser.MapHistoryFromDescriptor (FileParser.cs:129-150)."""
    out = []
    for frame in text.split(FRAME_SEP):
        lines = [l for l in frame.split("\n") if l]
        if not lines:
            continue
        t = float(lines[0])
        comps = [parse_gaussian(l) for l in lines[1:]]
        for _, m, _ in comps:
            if len(m) != dim:
                raise ValueError("wrong gaussian dimension")
        w = np.array([c[0] for c in comps])
        m = np.array([c[1] for c in comps]).reshape(-1, dim)
        P = np.array([c[2] for c in comps]).reshape(-1, dim, dim)
        out.append((t, (w, m, P)))
    return out


def trajectory_history_to_text(history, lossless=False):
    """Simulation.SerializedEstimate (Simulation.cs:172-181)."""
    return FRAME_SEP.join(g6(t) + "\n" + timed_array_to_text(rows, lossless) for t, rows in history)


def parse_trajectory_history(text, dim):
    """FileParser.TrajectoryHistoryFromDescriptor (FileParser.cs:66-95, smooth form)."""
    out = []
    for frame in text.split(FRAME_SEP):
        lines = [l for l in frame.split("\n") if l]
        if not lines:
            continue
        out.append((float(lines[0]), parse_timed_array(lines[1:], dim)))
    return out


def measurements_to_text(history):
    """Simulation.SerializedMeasurements (Simulation.cs:186-192): components in full precision."""
    return "\n".join(g6(t) + ":" + ";".join(" ".join(repr(float(v)) for v in z) for z in zs) for t, zs in history)


def parse_measurements(text, dim=3):
    """FileParser.MeasurementsFromDescriptor (FileParser.cs:176-228)."""
    out = []
    for frame in text.split("\n"):
        parts = frame.split(":")
        if len(parts) != 2:
            raise ValueError("bad measurement format: no ':' delimiter found")
        t = float(parts[0])
        zs = []
        for point in parts[1].split(";"):
            if point == "":
                continue
            comps = [float(c) for c in point.split(" ")]
            if len(comps) != dim:
                raise ValueError("wrong measurement dimension")
            zs.append(comps)
        out.append((t, np.array(zs, dtype=np.float64).reshape(-1, dim)))
    return out


# ------------------------------------------------------------------ the archive
@dataclass
class Recording:
    pose0: np.ndarray                      # scene: initial pose (7), measurer parameters (7), landmarks (n, 3)
    measurer: np.ndarray
    landmarks: np.ndarray
    trajectory: list = field(default_factory=list)     # [(t, state7)] groundtruth
    odometry: list = field(default_factory=list)       # [(t, reading6)]
    measurements: list = field(default_factory=list)   # [(t, z[m,3])]
    estimate: list = field(default_factory=list)       # [(t, [(t, state7)])] best particle's trajectory per frame
    maps: list = field(default_factory=list)           # [(t, (w, m, P))] best particle's map per frame
    vismaps: list = field(default_factory=list)
    tags: list = field(default_factory=list)           # [(t, message)]
    config: str = ""


def save(rec, filename, lossless=False):
    """Simulation.SaveToFile (Simulation.cs:391-488) without the sidebar video."""
    members = {
        "scene.world": scene_to_text(rec.pose0, rec.measurer, rec.landmarks, lossless),
        "trajectory.out": timed_array_to_text(rec.trajectory, lossless),
        "odometry.out": timed_array_to_text(rec.odometry, lossless),
        "estimate.out": trajectory_history_to_text(rec.estimate, lossless),
        "maps.out": map_history_to_text(rec.maps, lossless),
        "vismaps.out": map_history_to_text(rec.vismaps, lossless),
        "measurements.out": measurements_to_text(rec.measurements),
        "tags.out": "\n".join(g6(t) + " " + msg for t, msg in rec.tags),
        "config.cfg": rec.config,
    }
    with zipfile.ZipFile(filename, "w", zipfile.ZIP_DEFLATED) as zf:
        for name, text in members.items():
            zf.writestr(name, text)


def load(filename, extrainfo=True):
    """RecordVehicle.FromFile (RecordVehicle.cs:244-347): scene, trajectory, odometry and measurements are
    mandatory; estimate only when extrainfo; vismaps and tags optional."""
    with zipfile.ZipFile(filename) as zf:
        names = set(zf.namelist())

        def text(name):
            return io.TextIOWrapper(zf.open(name), encoding="utf-8", newline="").read()

        for need, what in (("scene.world", "scene"), ("trajectory.out", "trajectory"), ("odometry.out", "odometry"),
                           ("measurements.out", "measurement")):
            if need not in names:
                raise ValueError("Missing %s file" % what)
        if extrainfo and "estimate.out" not in names:
            raise ValueError("Missing estimate file")
        pose, measurer, landmarks = parse_scene(text("scene.world"))
        rec = Recording(pose, measurer, landmarks)
        rec.trajectory = parse_timed_array(text("trajectory.out").split("\n"), 7)
        rec.odometry = parse_timed_array(text("odometry.out").split("\n"), 6)
        rec.measurements = parse_measurements(text("measurements.out"), 3)
        if "estimate.out" in names:
            rec.estimate = parse_trajectory_history(text("estimate.out"), 7)
        if "maps.out" in names:
            rec.maps = parse_map_history(text("maps.out"), 3)
        if "vismaps.out" in names:
            rec.vismaps = parse_map_history(text("vismaps.out"), 3)
        if "tags.out" in names:
            for line in text("tags.out").split("\n"):
                if line.strip():
                    t, msg = line.split(" ", 1)
                    rec.tags.append((float(t), msg))
        if "config.cfg" in names:
            rec.config = text("config.cfg")
    return rec
