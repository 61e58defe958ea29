"""
TEST INFRASTRUCTURE ONLY -- imports the UNMODIFIED reference from /root/reference (this container only).

Purpose: (1) validate the CPU restatement in oracle/ (oracle.c / oracle.py) against the real reference,
(2) generate the golden fixtures committed under tests/golden/ (see tools/make_golden.py).
Nothing in the product package, the `-m gpu` tests, smoke() or bench.py may import this module:
/root/reference does not exist on the GPU box.

Shims (in memory, no reference file is modified), per SURVEY.md Appendix C:
  * matplotlib / matplotlib.pyplot stubbed (module-level imports at tools/snowfall/sampling.py:17,
    tools/wet_ground/augmentation.py:7, tools/wet_ground/phy_equations.py:9, tools/wet_ground/utils.py:8)
  * RANSACRegressor(loss='squared_loss') -> 'squared_error'   (tools/wet_ground/planes.py:35, sklearn>=1.2)
  * xedges[idx1] -> xedges[idx1[0]]                            (tools/wet_ground/augmentation.py:240-241, numpy>=1.23)
"""
import sys
import types
import inspect
import os

REF_ROOT = '/root/reference'


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, 'tools', 'snowfall'))


_loaded = None


def load():
    """Return a namespace with the reference modules (simulation, geometry, sampling, wet augmentation, ...)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError('reference tree not present (expected in the build container only)')
    sys.dont_write_bytecode = True
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    for name in ('matplotlib', 'matplotlib.pyplot', 'cv2_stub_unused'):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']

    import sklearn.linear_model as sklm
    import tools.wet_ground.planes as planes

    def _ransac_factory(loss='absolute_error', **k):
        return sklm.RANSACRegressor(loss='squared_error' if loss == 'squared_loss' else loss, **k)

    planes.RANSACRegressor = _ransac_factory

    import tools.wet_ground.augmentation as wet_aug
    src = inspect.getsource(wet_aug.estimate_laser_parameters)
    assert 'xedges[idx1]' in src
    src = src.replace('xedges[idx1]', 'xedges[idx1[0]]')
    ns = wet_aug.__dict__
    exec(compile(src, '<shimmed estimate_laser_parameters>', 'exec'), ns)

    import tools.snowfall.simulation as sim
    sim.estimate_laser_parameters = wet_aug.estimate_laser_parameters
    import tools.snowfall.geometry as geom
    import tools.snowfall.sampling as sampling
    import tools.wet_ground.phy_equations as phy

    _loaded = types.SimpleNamespace(sim=sim, geom=geom, sampling=sampling, wet_aug=wet_aug, planes=planes, phy=phy)
    return _loaded
