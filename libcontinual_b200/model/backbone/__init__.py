from .resnet import CifarResNet, cifar_resnet20, cifar_resnet32  # noqa: F401
