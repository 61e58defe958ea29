"""GPU test of the DenseDataset-style on-the-fly hook (lib/OpenPCDet/pcdet/datasets/dense/dense_dataset.py:749-837)."""
import numpy as np
import pytest

from lidar_snow_sim_b200.integrations.dense import OnTheFlyWeather
from lidar_snow_sim_b200.synthetic import synthetic_cloud

pytestmark = pytest.mark.gpu


def test_on_the_fly_weather(engine):
    pc = synthetic_cloud(seed=12, n_azimuth=1024)
    # evaluation mode / no keys: untouched
    assert OnTheFlyWeather({'SNOW': 'uniform_gunn_8in9'}, engine=engine)(pc, training=False) is pc
    assert OnTheFlyWeather({}, engine=engine)(pc) is pc
    np.random.seed(0)
    aug = OnTheFlyWeather({'SNOW': 'uniform_gunn_8in9', 'WET_SURFACE': '1in2', 'COUPLED': True}, engine=engine)
    outs = [aug(pc) for _ in range(6)]
    applied = [o for o in outs if o is not pc]
    assert len(applied) >= 3                                    # 8 in 9 chance
    for o in applied:
        assert o.shape[1] == 5 and 0 < o.shape[0] < pc.shape[0]
        assert set(np.unique(o[:, 4])) <= {0.0, 1.0, 2.0}
    assert len(aug._tables) >= 1                                # tables drawn once per (mode, rain rate), then cached
    # wet ground only
    np.random.seed(1)
    wet = OnTheFlyWeather({'WET_SURFACE': '1in2_norm'}, engine=engine)
    outs = [wet(pc) for _ in range(6)]
    assert any(o is not pc and o.dtype == np.float64 for o in outs)
