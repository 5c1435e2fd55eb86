"""B200-native Gibbs sweep engine behind the API of ExtendedRtIrtModeling.jl (hot path only).

The compute path is the CUDA shared library liberirt_b200.so (csrc/, C ABI in include/erirt_b200.h); this
package is the host-side mirror of the reference's interface for that path.  No CPU fallback exists."""
from . import _lib  # noqa: F401
from .api import (GibbsMlIrt, GibbsRtIrt, GibbsRtIrtCross, GibbsRtIrtCrossQr, GibbsRtIrtLatent,  # noqa: F401
                  GibbsRtIrtLatentQr, GibbsRtIrtNull, GibbsRtIrtQuantile, coef, getDic, getLogLikelihood, precis, sample,
                  sample_bang, sampleIndependentChains)
from .engine import Engine, ErirtError, k_nu_person, k_pg, k_philox, nccl_unique_id, trim_pool  # noqa: F401
from .simulate import (DeviceData, getBias, getRmse, setDataOnDevice, setDataMlIrt, setDataRtIrt, setDataRtIrtCross, setDataRtIrtLatent,  # noqa: F401
                       setDataRtIrtNull, setTrueParaMlIrt, setTrueParaRtIrt, setTrueParaRtIrtCross,
                       setTrueParaRtIrtLatent)
from .structs import InputData, InputData4R, InputPara, OutputDic, OutputPost, SimConditions, setCond  # noqa: F401
from .simtools import checkConvergence, comparePara, getMetrics, getMetrics2, runSimulation  # noqa: F401
from .ingest import readCsvData  # noqa: F401
