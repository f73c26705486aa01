"""openPMD checkpoint of the frame store and the fields: the layout contract of the
reference's openPMD plugin, so that a PIConGPU build can restart from / be diffed against our state.

What the reference writes (and reads back on `--checkpoint.restart`):
  * series attributes: openPMD 1.1.0 + ED-PIC extension, `picongpuIOVersionMajor/Minor`, software
    (include/picongpu/plugins/common/openPMDWriteMeta.hpp:111-137, openPMDVersion.def:47-57);
  * iteration attributes: dt, time, timeUnitSI, `iteration`, `sim_slides`, cell_width/height/depth, unit_*, mue0, eps0
    (openPMDWriteMeta.hpp:143-245; `iteration` and `sim_slides` are what the restart reads first,
    plugins/openPMD/openPMDWriter.x.cpp:1292-1298);
  * meshes E and B (FileCheckpointFields), components x, y, z, the whole global domain without guards as F[z][y][x]
    (writeField / writeFieldAttributes, openPMDWriter.x.cpp:1425-1488,1490-1665): unitDimension, timeOffset, geometry
    "cartesian", dataOrder "C", axisLabels (z, y, x), gridSpacing, gridGlobalOffset, gridUnitSI, fieldSmoothing; per
    component `position` (the Yee in-cell position, in x, y, z order as the reference passes it) and unitSI;
  * per species (WriteSpecies.hpp:205-275, restart/LoadSpecies.hpp:60-165): the frame attributes minus `multiMask` and
    `localCellIdx`, plus `totalCellIdx` -- i.e. records position (in-cell, float_X), positionOffset (= totalCellIdx,
    int32), momentum, weighting, constant records mass and charge, with unitSI / unitDimension / macroWeighted /
    weightingPower / timeOffset from traits/PICToOpenPMD.tpp:36-128 and unitless/speciesAttributes.unitless:60-240;
    species attributes particleShape, currentDeposition, particlePush, particleInterpolation, particleSmoothing;
  * particlePatches numParticles, numParticlesOffset, offset/{x,y,z}, extent/{x,y,z}: one patch per rank; a restarting
    rank finds ITS particles by comparing offset and extent with its local domain (LoadSpecies.hpp:190-255).

Container: neither HDF5 nor ADIOS2 (nor openPMD-api) exists in this image, so the tree is serialised the way
openPMD-api's JSON backend lays it out (one file per iteration, groups = objects with an "attributes" member, datasets
= objects with "datatype" and nested-list "data", attributes = {"datatype", "value"}); PIConGPU selects that backend
with `--checkpoint.openPMD.ext json`.  The tree builder (`checkpoint_tree`) is container independent.  Compatibility
with openPMD-api itself cannot be checked here ("parity unpinned" for the container); what the tests pin is the record
/ attribute contract above and a bit-exact dump -> restore round trip.  JSON is text: this is for small states (a diff
against the reference on a box where it builds), not for production checkpoints.

Not covered: PML auxiliary fields (the C ABI has no access to psi), RNG states and particle ids (none on this path).
"""
import json
import os

import numpy as np

from . import param as prm

OPENPMD_VERSION = "1.1.0"
PICONGPU_IO_VERSION = (3, 0)  # plugins/common/openPMDVersion.def:47,57
FILE_PATTERN = "checkpoint_%T"

_DT = {np.dtype(np.float32): "FLOAT", np.dtype(np.float64): "DOUBLE", np.dtype(np.int32): "INT", np.dtype(np.uint32): "UINT",
       np.dtype(np.uint64): "ULONG", np.dtype(np.int64): "LONG"}
_NP = {v: k for k, v in _DT.items()}
_BYTE_WIDTHS = {"BOOL": 1, "CHAR": 1, "DOUBLE": 8, "FLOAT": 4, "INT": 4, "LONG": 8, "LONGLONG": 8, "LONG_DOUBLE": 16, "SHORT": 2, "UCHAR": 1,
                "UINT": 4, "ULONG": 8, "ULONGLONG": 8, "USHORT": 2}

# seven openPMD base dimensions L, M, T, I, theta, N, J (traits/SIBaseUnits.hpp:34-44)
_DIM = dict(
    position=(1, 0, 0, 0, 0, 0, 0), positionOffset=(1, 0, 0, 0, 0, 0, 0), momentum=(1, 1, -1, 0, 0, 0, 0), weighting=(0,) * 7,
    mass=(0, 1, 0, 0, 0, 0, 0), charge=(0, 0, 1, 1, 0, 0, 0),
    E=(1, 1, -3, -1, 0, 0, 0),  # V / m (fields/FieldE.x.cpp:53-64)
    B=(0, 1, -2, -1, 0, 0, 0),  # T (fields/FieldB.x.cpp)
)
# Yee in-cell positions of the components x, y, z (fields/YeeCell.hpp:50-135)
_YEE = dict(E=((0.5, 0.0, 0.0), (0.0, 0.5, 0.0), (0.0, 0.0, 0.5)), B=((0.0, 0.5, 0.5), (0.5, 0.0, 0.5), (0.5, 0.5, 0.0)))
_SHAPE_ORDER = {prm.SHAPE_NGP: 0.0, prm.SHAPE_CIC: 1.0, prm.SHAPE_TSC: 2.0, prm.SHAPE_PQS: 3.0, prm.SHAPE_PCS: 4.0}
_PUSHER = {prm.PUSHER_BORIS: "Boris", prm.PUSHER_VAY: "Vay", prm.PUSHER_HIGUERA_CARY: "HigueraCary"}
_CURRENT = {prm.CURRENT_ESIRKEPOV: "Esirkepov", prm.CURRENT_EMZ: "EmZ"}
_SOLVER = {prm.SOLVER_YEE: "Yee", prm.SOLVER_LEHE: "Lehe"}


# ---- tree helpers ----------------------------------------------------------------------------------------------------
def _attr(value):
    """one attribute in the JSON backend's long form {"datatype", "value"}"""
    if isinstance(value, str):
        return {"datatype": "STRING", "value": value}
    if isinstance(value, (list, tuple)) and value and isinstance(value[0], str):
        return {"datatype": "VEC_STRING", "value": list(value)}
    a = np.asarray(value)
    if a.dtype == np.bool_:
        return {"datatype": "BOOL", "value": bool(a)}
    name = _DT[a.dtype]
    if a.ndim == 0:
        return {"datatype": name, "value": a.item()}
    return {"datatype": "VEC_" + name, "value": a.tolist()}


def _group(**attributes):
    return {"attributes": {k: _attr(v) for k, v in attributes.items()}}


def _dataset(array, **attributes):
    array = np.ascontiguousarray(array)
    d = _group(**attributes)
    d["datatype"] = _DT[array.dtype]
    d["data"] = array
    return d


def _constant(value, n, **attributes):
    """constant record component (RecordComponent::makeConstant): no dataset, attributes `value` and `shape`"""
    return _group(value=np.float64(value), shape=np.array([n], np.uint64), **attributes)


def _unit_dimension(key):
    return {"datatype": "ARR_DBL_7", "value": [float(v) for v in _DIM[key]]}


def _value(attr):
    return attr["value"]


# ---- what one rank contributes -----------------------------------------------------------------------------------------
def collect(sim):
    """Rank-local part of a checkpoint: interior E and B and, per species, the frame attributes with localCellIdx turned
    into totalCellIdx (local cell + offset of the local domain in the total domain; after window slides the total
    domain has moved on by `slides` local domains, Selection/MovingWindow semantics of WriteSpecies.hpp:300-330)."""
    from . import picstep

    p = sim.p
    g, n = p.guard_cells, p.grid
    inner = (slice(None), slice(g[2], g[2] + n[2]), slice(g[1], g[1] + n[1]), slice(g[0], g[0] + n[0]))
    slides = int(getattr(sim, "slides", 0))
    offset = [n[d] * p.rank_pos[d] for d in range(3)]
    total_offset = list(offset)
    total_offset[1] += slides * n[1]
    out = dict(grid=tuple(n), rank_pos=tuple(p.rank_pos), devices=tuple(p.devices), offset=tuple(offset), total_offset=tuple(total_offset),
               slides=slides, step=int(sim.step_index),
               E=np.ascontiguousarray(sim.download_field(picstep.FIELD_E)[inner]), B=np.ascontiguousarray(sim.download_field(picstep.FIELD_B)[inner]), species=[])
    for sp in p.species:
        pos, mom, w, cell = sim.download_particles(sp.name)
        cell = cell.astype(np.int64)
        tci = np.stack([cell % n[0] + total_offset[0], (cell // n[0]) % n[1] + total_offset[1], cell // (n[0] * n[1]) + total_offset[2]]).astype(np.int32)
        out["species"].append(dict(name=sp.name, position=np.ascontiguousarray(pos, np.float32), positionOffset=tci,
                                   momentum=np.ascontiguousarray(mom, np.float32), weighting=np.ascontiguousarray(w, np.float32)))
    return out


# ---- the openPMD tree ------------------------------------------------------------------------------------------------------
def checkpoint_tree(p, parts):
    """The openPMD hierarchy of one iteration from the parts of all ranks (`collect`), ranks in any order.
    `p`: SimParams of any rank (units, cell sizes, policies)."""
    parts = sorted(parts, key=lambda q: q["rank_pos"][::-1])
    step, slides = parts[0]["step"], parts[0]["slides"]
    if any(q["step"] != step or q["slides"] != slides for q in parts):
        raise ValueError("the parts belong to different steps")
    gg = p.global_grid
    cell = np.array(p.cell_size, np.float32)
    unit_length = float(p.unit_length)
    cell_si = [float(cell[d]) * unit_length for d in range(3)]

    root = _group(openPMD=OPENPMD_VERSION, openPMDextension=np.uint32(1), basePath="/data/%T/", meshesPath="fields/", particlesPath="particles/",
                  iterationEncoding="fileBased", iterationFormat=FILE_PATTERN, software="picstep-b200 (PIConGPU checkpoint layout)",
                  softwareVersion="0.2", picongpuIOVersionMajor=np.int32(PICONGPU_IO_VERSION[0]), picongpuIOVersionMinor=np.int32(PICONGPU_IO_VERSION[1]))
    it = _group(dt=np.float32(p.dt), time=np.float32(np.float32(step) * np.float32(p.dt)), timeUnitSI=np.float64(p.unit_time),
                iteration=np.uint32(step), sim_slides=np.uint32(slides), cell_width=cell[0], cell_height=cell[1], cell_depth=cell[2],
                unit_energy=np.float64(p.unit_energy), unit_length=np.float64(p.unit_length), unit_speed=np.float64(p.unit_speed),
                unit_time=np.float64(p.unit_time), unit_mass=np.float64(p.unit_mass), unit_charge=np.float64(p.unit_charge),
                unit_efield=np.float64(p.unit_efield), unit_bfield=np.float64(p.unit_bfield), mue0=np.float32(p.mue0), eps0=np.float32(p.eps0))
    root["data"] = {str(step): it}

    # -- meshes ---------------------------------------------------------------------------------------------------------
    solver = _SOLVER[p.field_solver]
    fields = it["fields"] = _group(fieldSolver=solver, currentSmoothing="Binomial" if p.current_interpolation else "none", chargeCorrection="none")
    for name, unit in (("E", p.unit_efield), ("B", p.unit_bfield)):
        full = np.zeros((3, gg[2], gg[1], gg[0]), np.float32)
        for q in parts:
            o, n = q["offset"], q["grid"]
            full[:, o[2]:o[2] + n[2], o[1]:o[1] + n[1], o[0]:o[0] + n[0]] = q[name]
        # gridGlobalOffset: cell size times (window offset + slide offset), F[z][y][x] order
        goff = [0.0, float(np.float64(cell[1]) * np.float64(slides * p.grid[1])), 0.0]
        mesh = fields[name] = _group(timeOffset=np.float32(0.0), geometry="cartesian", dataOrder="C", axisLabels=["z", "y", "x"],
                                     gridSpacing=cell[::-1].copy(), gridGlobalOffset=np.array(goff, np.float64), gridUnitSI=np.float64(unit_length),
                                     fieldSmoothing="none")
        mesh["attributes"]["unitDimension"] = _unit_dimension(name)
        for c, comp in enumerate("xyz"):
            mesh[comp] = _dataset(full[c], position=np.array(_YEE[name][c], np.float32), unitSI=np.float64(unit))

    # -- species ---------------------------------------------------------------------------------------------------------
    particles = it["particles"] = {}
    for s, sp in enumerate(p.species):
        per = [q["species"][s] for q in parts]
        counts = np.array([x["weighting"].shape[0] for x in per], np.uint64)
        ntot = int(counts.sum())
        grp = particles[sp.name] = _group(particleShape=np.float64(_SHAPE_ORDER[p.shape]), currentDeposition=_CURRENT[p.current_solver],
                                          particlePush=_PUSHER[p.pusher], particleInterpolation="uniform", particleSmoothing="none")

        def record(key, macro_weighted, weighting_power):
            r = _group(macroWeighted=np.int32(macro_weighted), weightingPower=np.float64(weighting_power), timeOffset=np.float64(0.0))
            r["attributes"]["unitDimension"] = _unit_dimension(key)
            return r

        # speciesAttributes.unitless: position (in cell) and totalCellIdx scale with the cell size; momentum with mass * speed
        rec = grp["position"] = record("position", 0, 0.0)
        off = grp["positionOffset"] = record("positionOffset", 0, 0.0)
        mom = grp["momentum"] = record("momentum", 1, 1.0)
        for c, comp in enumerate("xyz"):
            rec[comp] = _dataset(np.concatenate([x["position"][c] for x in per]) if per else np.zeros(0, np.float32), unitSI=np.float64(cell_si[c]))
            off[comp] = _dataset(np.concatenate([x["positionOffset"][c] for x in per]), unitSI=np.float64(cell_si[c]))
            mom[comp] = _dataset(np.concatenate([x["momentum"][c] for x in per]), unitSI=np.float64(p.unit_mass * p.unit_speed))
        wrec = grp["weighting"] = record("weighting", 1, 1.0)
        wrec.update({k: v for k, v in _dataset(np.concatenate([x["weighting"] for x in per]), unitSI=np.float64(1.0)).items() if k != "attributes"})
        wrec["attributes"]["unitSI"] = _attr(np.float64(1.0))
        # constant records: mass and charge of ONE real particle in PIC units (GetMassOrZero / GetChargeOrZero), unitSI = unit
        for key, value, unit in (("mass", float(np.float32(p.base_mass) * np.float32(sp.mass_ratio)), p.unit_mass),
                                 ("charge", float(np.float32(p.base_charge) * np.float32(sp.charge_ratio)), p.unit_charge)):
            r = grp[key] = record(key, 0, 1.0)
            c = _constant(value, ntot, unitSI=np.float64(unit))
            r["attributes"].update(c["attributes"])
        # particle patches: one per rank, in units of cells of the total domain
        patches = grp["particlePatches"] = {"attributes": {}}
        patches["numParticles"] = _dataset(counts, unitSI=np.float64(1.0))
        patches["numParticlesOffset"] = _dataset(np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.uint64), unitSI=np.float64(1.0))
        patches["offset"], patches["extent"] = {"attributes": {}}, {"attributes": {}}
        patches["offset"]["attributes"]["unitDimension"] = _unit_dimension("position")
        patches["extent"]["attributes"]["unitDimension"] = _unit_dimension("position")
        for c, comp in enumerate("xyz"):
            patches["offset"][comp] = _dataset(np.array([q["total_offset"][c] for q in parts], np.uint64), unitSI=np.float64(cell_si[c]))
            patches["extent"][comp] = _dataset(np.array([q["grid"][c] for q in parts], np.uint64), unitSI=np.float64(cell_si[c]))
    root["platform_byte_widths"] = dict(_BYTE_WIDTHS)
    return root


# ---- JSON container (openPMD-api JSON backend layout) --------------------------------------------------------------------------
def _to_jsonable(node):
    if isinstance(node, dict):
        return {k: _to_jsonable(v) for k, v in node.items()}
    if isinstance(node, np.ndarray):
        return node.tolist()
    if isinstance(node, np.generic):
        return node.item()
    return node


def _from_json(node):
    if isinstance(node, dict):
        out = {k: _from_json(v) for k, v in node.items() if k != "data" or "datatype" not in node}
        if "datatype" in node and "data" in node:
            out["data"] = np.array(node["data"], dtype=_NP[node["datatype"]])
        return out
    return node


def file_name(directory, step):
    return os.path.join(directory, FILE_PATTERN.replace("%T", str(step)) + ".json")


def write(p, parts, directory):
    """Write iteration `step` of the checkpoint series from the parts of all ranks; returns the file name."""
    tree = checkpoint_tree(p, parts)
    os.makedirs(directory, exist_ok=True)
    name = file_name(directory, parts[0]["step"])
    with open(name, "w") as f:
        json.dump(_to_jsonable(tree), f)
    return name


def write_distributed(p, part, directory, rank, world):
    """One process per rank (torch.distributed is initialised): the parts are gathered on rank 0, which writes the
    file; every rank returns the file name after a barrier."""
    import torch.distributed as dist

    parts = [None] * world if rank == 0 else None
    dist.gather_object(part, parts, dst=0)
    if rank == 0:
        write(p, parts, directory)
    dist.barrier()
    return file_name(directory, part["step"])


def read(name):
    with open(name) as f:
        return _from_json(json.load(f))


# ---- restart ---------------------------------------------------------------------------------------------------------------
def check_restart_compatibility(tree):
    """checkIOFileVersionRestartCompatibility (openPMDWriter.x.cpp:1196-1270): same major file format version"""
    a = tree["attributes"]
    major = _value(a["picongpuIOVersionMajor"]) if "picongpuIOVersionMajor" in a else 0
    if major != PICONGPU_IO_VERSION[0]:
        raise ValueError("checkpoint has picongpuIOVersionMajor %s, this reader handles %d" % (major, PICONGPU_IO_VERSION[0]))


def rank_part(tree, p, step=None):
    """The part of the checkpoint that belongs to the rank described by `p` (grid, rank_pos): its block of E and B, and
    for every species the particles of the patch whose offset and extent equal the local domain (getPatchIdx,
    LoadSpecies.hpp:190-255) with totalCellIdx turned back into the local linear cell index."""
    check_restart_compatibility(tree)
    steps = sorted(tree["data"], key=int)
    key = str(step) if step is not None else steps[-1]
    it = tree["data"][key]
    at = it["attributes"]
    if _value(at["iteration"]) != int(key):
        raise ValueError("iteration attribute does not match the iteration key")
    slides = int(_value(at["sim_slides"]))
    n = p.grid
    offset = [n[d] * p.rank_pos[d] for d in range(3)]
    total_offset = list(offset)
    total_offset[1] += slides * n[1]
    out = dict(step=int(key), slides=slides, species=[])
    for name in ("E", "B"):
        mesh = it["fields"][name]
        blk = [mesh[c]["data"][offset[2]:offset[2] + n[2], offset[1]:offset[1] + n[1], offset[0]:offset[0] + n[0]] for c in "xyz"]
        if blk[0].shape != (n[2], n[1], n[0]):
            raise ValueError("the local domain is not inside the mesh of the checkpoint")
        out[name] = np.ascontiguousarray(np.stack(blk), np.float32)
    for sp in p.species:
        grp = it["particles"][sp.name]
        pt = grp["particlePatches"]
        offs = np.stack([pt["offset"][c]["data"] for c in "xyz"], axis=1).astype(np.int64)
        exts = np.stack([pt["extent"][c]["data"] for c in "xyz"], axis=1).astype(np.int64)
        hit = [i for i in range(offs.shape[0]) if tuple(offs[i]) == tuple(total_offset) and tuple(exts[i]) == tuple(n)]
        if not hit:
            raise ValueError("Error while restarting: no particle patch matches the required offset and extent")
        i = hit[0]
        a = int(pt["numParticlesOffset"]["data"][i])
        b = a + int(pt["numParticles"]["data"][i])
        tci = np.stack([grp["positionOffset"][c]["data"][a:b] for c in "xyz"]).astype(np.int64)
        loc = tci - np.array(total_offset, np.int64)[:, None]
        if loc.size and (loc.min() < 0 or (loc >= np.array(n)[:, None]).any()):
            raise ValueError("particle outside of its patch")
        out["species"].append(dict(name=sp.name, position=np.ascontiguousarray(np.stack([grp["position"][c]["data"][a:b] for c in "xyz"]), np.float32),
                                   momentum=np.ascontiguousarray(np.stack([grp["momentum"][c]["data"][a:b] for c in "xyz"]), np.float32),
                                   weighting=np.ascontiguousarray(grp["weighting"]["data"][a:b], np.float32),
                                   cell=(loc[0] + n[0] * (loc[1] + n[1] * loc[2])).astype(np.int32)))
    return out


def restore(sim, tree, step=None):
    """Simulation::init with a restart step (Simulation.hpp:436-470 + openPMDWriter::doRestart, openPMDWriter.x.cpp:1272-1340):
    window slides first, then fields, then particles; guards of E and B are filled by an exchange as at the end of a step.
    `sim` is a freshly created Simulation of this rank (after comm_init when there are several)."""
    from . import picstep

    part = rank_part(tree, sim.p, step)
    p = sim.p
    if getattr(p, "absorber_kind", 0) == 2:
        raise NotImplementedError("the PML auxiliary fields are not part of this checkpoint")
    for _ in range(part["slides"] - int(getattr(sim, "slides", 0))):
        sim.slide()
        part = rank_part(tree, sim.p, step)  # the rank has moved inside the window
    g, n, N = p.guard_cells, p.grid, p.padded
    for fld, name in ((picstep.FIELD_E, "E"), (picstep.FIELD_B, "B")):
        full = np.zeros((3, N[2], N[1], N[0]), np.float32)
        full[:, g[2]:g[2] + n[2], g[1]:g[1] + n[1], g[0]:g[0] + n[0]] = part[name]
        sim.upload_field(fld, full)
    for s in part["species"]:
        sim.upload_particles(s["name"], s["position"], s["momentum"], s["weighting"], s["cell"])
    sim.step_index = part["step"]
    sim.field_exchange(picstep.FIELD_E)
    sim.field_exchange(picstep.FIELD_B)
    return part
