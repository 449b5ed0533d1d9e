from .resnet import CifarResNet, ResNet18, cifar_resnet20, cifar_resnet32, resnet18, resnet32_V2  # noqa: F401
