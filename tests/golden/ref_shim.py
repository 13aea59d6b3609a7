"""Import shim for the UNMODIFIED reference at /root/reference (test infrastructure only).

Used only by tests/golden/make_golden.py, in the build container, to produce the
committed golden fixtures. /root/reference does not exist on the GPU box and nothing
in tests/, bench.py or the package imports this module at run time.

Recipe: SURVEY.md App. D.  Six third-party imports the reference needs but the image
lacks (easydict lmdb Bio torch_scatter omegaconf tree) are registered as stub modules,
and the hard-coded names.txt open at models_con/pep_dataloader.py:36-39 is answered
with an empty file while the import runs.
"""
import builtins
import importlib.machinery
import io
import sys
import types

REFERENCE_ROOT = "/root/reference"


class EasyDict(dict):
    """Attribute dict with nested conversion (stand-in for easydict.EasyDict)."""

    def __init__(self, d=None, **kw):
        super().__init__()
        d = dict(d or {})
        d.update(kw)
        for k, v in d.items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, EasyDict):
            v = EasyDict(v)
        elif isinstance(v, (list, tuple)):
            v = type(v)(EasyDict(x) if isinstance(x, dict) else x for x in v)
        dict.__setitem__(self, k, v)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    __setattr__ = __setitem__


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def _map_structure(fn, s):
    if isinstance(s, (list, tuple)):
        return type(s)(_map_structure(fn, x) for x in s)
    if isinstance(s, dict):
        return {k: _map_structure(fn, v) for k, v in s.items()}
    return fn(s)


class _Any:
    def __init__(self, *a, **k):
        pass

    def __getattr__(self, k):
        return _Any()

    def __call__(self, *a, **k):
        return _Any()


def _module_getattr(k):
    if k.startswith("__"):
        raise AttributeError(k)
    return _Any


_loaded = None


def load_reference():
    """Returns a namespace with the reference's FlowModel, load_config and friends."""
    global _loaded
    if _loaded is not None:
        return _loaded
    import torch, numpy, scipy, pandas  # noqa: F401  real modules first
    try:
        import wandb  # noqa: F401
    except Exception:
        _stub("wandb")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _stub("easydict", EasyDict=EasyDict)
    _stub("lmdb")
    _stub("omegaconf", OmegaConf=object)
    _stub("torch_scatter", scatter_add=None, scatter=None)
    _stub("tree", map_structure=_map_structure)
    bio = _stub("Bio", BiopythonWarning=Warning)
    for sub in ["PDB", "PDB.Chain", "PDB.Residue", "PDB.PDBParser", "PDB.MMCIFParser",
                "PDB.StructureBuilder", "PDB.PDBExceptions", "PDB.Selection", "PDB.PDBIO",
                "PDB.DSSP", "PDB.Polypeptide", "SeqUtils", "Seq", "SeqRecord", "SeqIO"]:
        _stub("Bio." + sub).__getattr__ = _module_getattr

    class PDBConstructionException(Exception):
        pass

    sys.modules["Bio.PDB.PDBExceptions"].PDBConstructionException = PDBConstructionException
    sys.modules["Bio.PDB"].PDBExceptions = sys.modules["Bio.PDB.PDBExceptions"]
    sys.modules["Bio.PDB"].Selection = sys.modules["Bio.PDB.Selection"]
    bio.PDB = sys.modules["Bio.PDB"]

    _open = builtins.open

    def _patched(path, *a, **k):
        if isinstance(path, str) and path.endswith("pepflowww/Data/names.txt"):
            return io.StringIO("")
        return _open(path, *a, **k)

    builtins.open = _patched
    try:
        from models_con.flow_model import FlowModel
        import models_con.flow_model as flow_model_mod
    finally:
        builtins.open = _open
    from pepflow.utils.misc import load_config
    ns = types.SimpleNamespace(FlowModel=FlowModel, flow_model_mod=flow_model_mod,
                               load_config=load_config)
    _loaded = ns
    return ns
