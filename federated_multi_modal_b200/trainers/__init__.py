from .maple import CustomCLIP, MaPLe, MultiModalPromptLearner, TextEncoder, load_clip_to_cpu  # noqa: F401
from .maple_fed import MaPLeFederated  # noqa: F401
from .client_datamanager import ClientDataManager, Datum, GpuAugment, sample_rrc_params  # noqa: F401
from .data_partition import partition_dataset_iid, partition_dataset_dirichlet  # noqa: F401
