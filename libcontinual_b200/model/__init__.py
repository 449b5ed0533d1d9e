"""Method plugins (mirror of core/model/__init__.py for the hot-path methods) and backbone factories."""
from .backbone import CifarResNet, ResNet18, cifar_resnet20, cifar_resnet32, resnet18, resnet32_V2  # noqa: F401
from .resnet_methods import EWC, LUCIR, LWF, Finetune, ICarl  # noqa: F401
from .l2p import L2P, ViTZoo, vit_pt_imnet  # noqa: F401
from .inflora import InfLoRA_OPT, SiNet  # noqa: F401
from .dualprompt import DualPrompt, DualPromptPool  # noqa: F401
from .codaprompt import CodaPrompt, CodaPromptPool  # noqa: F401
from .sd_lora import SD_LoRA  # noqa: F401
from .inflora_orig import InfLoRA, SiNet_vit  # noqa: F401
from .gpm import GPM, AlexNet_TRGP  # noqa: F401
