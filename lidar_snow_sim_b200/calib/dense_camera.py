"""
Camera calibration of the SeeingThroughFog (DENSE) HDL-64 rig used by the camera field-of-view filter.

Calibration DATA of the dataset (the values the reference ships as lib/OpenPCDet/data/dense/calib_hdl64.txt, lines
P2 / R0_rect / Tr_velo_to_cam), parsed to float32 exactly as lib/OpenPCDet/pcdet/utils/calibration_kitti.py:5-20
does (np.array(strings, dtype=np.float32)).  Consumed by tools/snowfall/simulation.py:532-540.
"""
import numpy as np

_P2 = ('2355.722801 0.0 988.138054 0.0 0.0 2355.722801 508.051838 0.0 0.0 0.0 1.0 0.0')
_R0 = ('1.0 0.0 0.0 0.0 1.0 0.0 0.0 0.0 1.0')
_V2C = ('0.008165559536184884 -0.9999654854673377 -0.0015334638999611712 0.08749686458036277 '
        '-0.007155442920435107 0.001475045689574084 -0.9999733115822772 -0.4164032939864676 '
        '0.9999410599000692 0.008176314223727743 -0.007143151380269915 -0.6955577980059064')


def parse_calib_text(path):
    """Read a KITTI-style calib file with P2 / R0_rect / Tr_velo_to_cam on lines 3, 5, 6 (calibration_kitti.py:5-20)."""
    with open(path) as f:
        lines = f.readlines()
    return {'P2': np.array(lines[2].strip().split(' ')[1:], dtype=np.float32).reshape(3, 4),
            'R0': np.array(lines[4].strip().split(' ')[1:], dtype=np.float32).reshape(3, 3),
            'V2C': np.array(lines[5].strip().split(' ')[1:], dtype=np.float32).reshape(3, 4),
            'img_shape': (1024, 1920)}


STF_HDL64_CAMERA = {
    'P2': np.array(_P2.split(' '), dtype=np.float32).reshape(3, 4),
    'R0': np.array(_R0.split(' '), dtype=np.float32).reshape(3, 3),
    'V2C': np.array(_V2C.split(' '), dtype=np.float32).reshape(3, 4),
    'img_shape': (1024, 1920),          # simulation.py:536
}
