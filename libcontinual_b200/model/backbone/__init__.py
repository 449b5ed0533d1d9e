from .resnet import CifarResNet, cifar_resnet20, cifar_resnet32, resnet32_V2  # noqa: F401
