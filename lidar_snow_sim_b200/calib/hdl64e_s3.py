"""
Velodyne HDL-64E S3 per-laser constants used by the snowfall received-power solve.

These are sensor calibration DATA (values of the public HDL-64E S3 calibration the reference ships as
calib/20171102_64E_S3.yaml, fields focal_distance / focal_slope / min_intensity / vert_correction), restated as a
plain table so the engine does not need the reference tree at run time.  Consumed where the reference reads them:
tools/snowfall/simulation.py:72-76 (min_intensity default 0 when absent, focal_distance, focal_slope) and :123-126
(max_intensity 230 for channels 53, 55, 56, 58).  `vert_correction` [rad] is only used by the synthetic cloud
generator.

A user-supplied YAML in the reference's format can be loaded with `load_sensor_yaml`.
"""
import numpy as np

# (laser_id, focal_distance, focal_slope, min_intensity or None, vert_correction [rad])
HDL64E_S3 = [
    (0, 8.0, 0.94999999, 40, -0.12658417533780103),
    (1, 12.5, 1.0, 40, -0.12175346699676513),
    (2, 24.0, 0.69999999, 40, 0.00356118725906175),
    (3, 24.0, 0.75, 40, 0.008262563386399281),
    (4, 18.0, 0.5, 40, -0.11611041958999802),
    (5, 24.0, 0.75, 40, -0.10965216205088012),
    (6, 13.0, 2.0, 40, -0.14884568763136827),
    (7, 24.0, 0.75, 40, -0.14504505690689615),
    (8, 23.5, 0.40000001, 40, -0.10318468393732395),
    (9, 19.5, 0.60000002, 40, -0.09772098481916251),
    (10, 6.5, 1.2, 40, -0.13923599552700858),
    (11, 21.0, 1.1, 40, -0.13261418667434482),
    (12, 24.0, 0.69999999, 40, -0.05525412215182704),
    (13, 24.0, 0.85000002, 40, -0.04975012163590781),
    (14, 14.0, 1.2, 20, -0.0910352183760594),
    (15, 24.0, 0.75, 40, -0.08495012793898118),
    (16, 24.0, 0.64999998, 40, -0.043427018903405855),
    (17, 24.0, 0.80000001, 40, -0.03812106264303967),
    (18, 24.0, 0.69999999, 40, -0.07926501984257026),
    (19, 24.0, 0.75, 40, -0.07438795730193064),
    (20, 24.0, 0.64999998, 40, -0.032200364768206785),
    (21, 24.0, 0.75, 20, -0.027094473854667452),
    (22, 24.0, 0.69999999, 10, -0.06808320147052566),
    (23, 24.0, 0.69999999, 40, -0.06218028850077059),
    (24, 11.0, 1.35, 10, 0.015211696348436898),
    (25, 21.0, 0.64999998, 40, 0.021137769837365927),
    (26, 24.0, 1.2, 10, -0.02014825541794774),
    (27, 24.0, 1.0, 40, -0.013608856125641236),
    (28, 24.0, 0.60000002, 20, 0.027062359796430756),
    (29, 24.0, 0.64999998, 10, 0.03257665775493676),
    (30, 24.0, 0.80000001, 40, -0.008703503368623128),
    (31, 24.0, 0.75, 10, -0.0019580150746423583),
    (32, 12.0, 2.0, 40, -0.39527185409579646),
    (33, 0.25, 0.94999999, 10, -0.3877619815771452),
    (34, 12.0, 0.40000001, None, -0.19930022938936873),
    (35, 9.5, 1.45, None, -0.18847487706375146),
    (36, 10.0, 1.55, None, -0.38038196155484233),
    (37, 9.5, 1.5, None, -0.3699409132772216),
    (38, 9.0, 1.35, None, -0.4331233680891581),
    (39, 0.25, 1.0, None, -0.42365160586492506),
    (40, 0.25, 0.80000001, None, -0.3610257714580346),
    (41, 12.0, 1.95, None, -0.3497067702067833),
    (42, 0.25, 1.05, None, -0.41563902616182186),
    (43, 8.0, 1.45, None, -0.4037726372972675),
    (44, 9.5, 1.5, None, -0.2870427654491517),
    (45, 9.5, 0.40000001, None, -0.2787493448161101),
    (46, 11.5, 2.0, None, -0.3424711063669128),
    (47, 0.25, 0.80000001, None, -0.3313646164053843),
    (48, 14.5, 0.5, None, -0.2715549929196343),
    (49, 11.5, 1.7, None, -0.2620447461853948),
    (50, 10.0, 1.5, None, -0.32697130851884926),
    (51, 14.5, 0.40000001, None, -0.3168539490043034),
    (52, 9.0, 0.40000001, None, -0.25363534842355817),
    (53, 12.5, 1.7, None, -0.2409525119882134),
    (54, 6.5, 0.40000001, None, -0.3075976651295216),
    (55, 12.0, 1.65, None, -0.29492149585191946),
    (56, 9.5, 1.6, None, -0.17958427457384749),
    (57, 0.25, 0.60000002, None, -0.1716573102518716),
    (58, 10.0, 1.9, None, -0.2343876828996395),
    (59, 0.25, 0.55000001, None, -0.22547306467922804),
    (60, 5.0, 0.80000001, None, -0.1647032640134567),
    (61, 0.25, 0.44999999, None, -0.15454155742794698),
    (62, 10.0, 1.1, None, -0.2184708037202268),
    (63, 15.0, 0.40000001, None, -0.2083180489284506),
]

NUM_LASERS = 64
MAX_INTENSITY_230_CHANNELS = (53, 55, 56, 58)      # tools/snowfall/simulation.py:123-126


def sensor_arrays(table=None):
    """Return float64 arrays (focal_distance, focal_slope, min_intensity, max_intensity), one entry per channel."""
    table = HDL64E_S3 if table is None else table
    n = len(table)
    fd = np.array([r[1] for r in table], dtype=np.float64)
    fs = np.array([r[2] for r in table], dtype=np.float64)
    mi = np.array([0.0 if r[3] is None else float(r[3]) for r in table], dtype=np.float64)
    mx = np.array([230.0 if c in MAX_INTENSITY_230_CHANNELS else 255.0 for c in range(n)], dtype=np.float64)
    return fd, fs, mi, mx


def vert_corrections(table=None):
    table = HDL64E_S3 if table is None else table
    return np.array([r[4] for r in table], dtype=np.float64)


def load_sensor_yaml(path):
    """Load a calibration file in the reference's YAML layout (`lasers: [{focal_distance, focal_slope, ...}]`)."""
    import yaml
    with open(path, 'r') as stream:
        d = yaml.safe_load(stream)
    return [(l.get('laser_id', i), l['focal_distance'], l['focal_slope'], l.get('min_intensity'),
             l.get('vert_correction', 0.0)) for i, l in enumerate(d['lasers'])]
