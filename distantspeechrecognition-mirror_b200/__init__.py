"""distantspeechrecognition-mirror_b200: B200-native (sm_100a) subband front end of the Beamforming Toolkit --
OverSampledDFTAnalysisBank -> SubbandDS / SubbandMVDR -> OverSampledDFTSynthesisBank -- behind the
reference's stream API.

The directory name carries a hyphen (it mirrors the upstream repository name); import it through the
``btk_b200`` shim at the repository root:  ``import btk_b200``.
"""
from . import _capi, sharding, streams, workloads  # noqa: F401
from ._capi import (BtkError, Plan, calc_all_delays, calc_delays_polar, design_analysis_nyquist,  # noqa: F401
                    design_analysis_prototype, design_synthesis_nyquist, design_synthesis_prototype, device_count, lib)
from .streams import (ChannelExtractionFeaturePtr, Conversion24bit2FloatPtr, IterativeSampleFeaturePtr,  # noqa: F401
                      OverSampledDFTAnalysisBankPtr, OverSampledDFTSynthesisBankPtr, SampleFeaturePtr,  # noqa: F401
                      SnapShotArrayPtr, SpectralMatrixArrayPtr, SubbandDSPtr, SubbandGSCPtr, SubbandMVDRPtr,
                      ZelinskiPostFilterPtr)

__all__ = ["Plan", "BtkError", "device_count", "lib", "design_analysis_prototype", "design_synthesis_prototype", "workloads", "streams", "SampleFeaturePtr",
           "OverSampledDFTAnalysisBankPtr", "OverSampledDFTSynthesisBankPtr", "SubbandDSPtr", "SubbandGSCPtr", "SubbandMVDRPtr", "ZelinskiPostFilterPtr",
           "SnapShotArrayPtr", "SpectralMatrixArrayPtr", "calc_delays_polar", "calc_all_delays"]
