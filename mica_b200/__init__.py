"""mica_b200 -- B200-native voxel-parallel map pipeline of MICA (cryo-EM + AlphaFold3).

Hot path only: resample -> normalise -> AF3 rasterise -> cube
extract -> [model] -> softmax/argmax + stitch, as hand-written sm_100a CUDA behind a
C ABI (include/mica_b200.h), with host-side mirrors of the reference's entry points:

    from mica_b200.preprocessing import DataPreprocessor      # utils/preprocessing.py
    from mica_b200.create_grids import GridCreator            # utils/create_grids.py
    from mica_b200.predict import CryoEMPredictor             # utils/predict.py

Importing a submodule that touches the GPU requires mica_b200/libmica_b200.so
(``python -m mica_b200.build``); there is no CPU fallback."""
__version__ = '0.1.0'
