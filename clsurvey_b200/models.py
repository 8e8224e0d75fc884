"""Host-side model definitions (kept in PyTorch, like the reference: src/models/{net,VGGSlim}.py).

The engine only needs the reference's module *structure*: `features` (Conv2d/ReLU/MaxPool2d Sequential), `avgpool`
(Identity | AdaptiveAvgPool2d) and `classifier` (Linear/ReLU/Dropout Sequential) with state_dict keys identical to
the reference's models, so pickled reference models and these are interchangeable.
"""
import torch
import torch.nn as nn

# channel configs of src/models/VGGSlim.py:13-24 ('M' = MaxPool2d(2,2))
cfg = {
    "19normal": [64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M", 512, 512, 512, 512, "M"],
    "16normal": [64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "M", 512, 512, 512, "M"],
    "11normal": [64, "M", 128, "M", 256, 256, "M", 512, 512, "M", 512, 512, "M"],
    "small_VGG9": [64, "M", 64, "M", 64, 64, "M", 128, 128, "M"],
    "base_VGG9": [64, "M", 64, "M", 128, 128, "M", 256, 256, "M"],
    "wide_VGG9": [64, "M", 128, "M", 256, 256, "M", 512, 512, "M"],
    "deep_VGG22": [64, "M", 64, 64, 64, 64, 64, 64, "M", 128, 128, 128, 128, 128, 128, "M",
                   256, 256, 256, 256, 256, 256, "M"],
}


def make_layers(config, in_channels=3):
    """Conv3x3(pad 1)+ReLU stacks with 2x2 max-pools (VGGSlim.py:27-40; batch-norm variants are out of scope)."""
    layers = []
    for v in config:
        if v == "M":
            layers.append(nn.MaxPool2d(kernel_size=2, stride=2))
        else:
            layers += [nn.Conv2d(in_channels, v, kernel_size=3, padding=1), nn.ReLU(inplace=True)]
            in_channels = v
    return nn.Sequential(*layers)


class VGGSlim(nn.Module):
    """Same attribute names / state_dict keys as the reference VGGSlim (VGGSlim.py:43-76)."""

    def __init__(self, config="11normal", num_classes=20, classifier_inputdim=512 * 2 * 2, classifier_dim1=512,
                 classifier_dim2=512, dropout=False, init_weights=True):
        super().__init__()
        self.features = make_layers(cfg[config] if isinstance(config, str) else config)
        self.avgpool = nn.Identity()
        if dropout:
            self.classifier = nn.Sequential(
                nn.Linear(classifier_inputdim, classifier_dim1), nn.ReLU(True), nn.Dropout(),
                nn.Linear(classifier_dim1, classifier_dim2), nn.ReLU(True), nn.Dropout(),
                nn.Linear(classifier_dim2, num_classes))
        else:
            self.classifier = nn.Sequential(
                nn.Linear(classifier_inputdim, classifier_dim1), nn.ReLU(True),
                nn.Linear(classifier_dim1, classifier_dim2), nn.ReLU(True),
                nn.Linear(classifier_dim2, num_classes))
        if init_weights:
            self._initialize_weights()

    def _initialize_weights(self):
        # torchvision VGG init: kaiming-normal(fan_out) convs, N(0, 0.01) linears, zero biases
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
                nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.Linear):
                nn.init.normal_(m.weight, 0, 0.01)
                nn.init.constant_(m.bias, 0)

    def forward(self, x):
        x = self.features(x)
        x = self.avgpool(x)
        x = torch.flatten(x, 1)
        return self.classifier(x)


def _n_pools(config):
    return sum(1 for v in config if v == "M")


def make_vgg(name, input_hw=(64, 64), num_classes=20):
    """`<config>_cl_<d1>_<d2>[_DROP]`, e.g. small_VGG9_cl_128_128, VGG11_cl_512_512 (net.py:15-36,243-262).

    'VGG11' is registered here (cfg['11normal'], 5 pools): the reference defines the config but never wires a name
    for it."""
    parts = name.split("_cl_")
    base = parts[0]
    key = "11normal" if base in ("VGG11", "11normal") else base
    if key not in cfg:
        raise NotImplementedError("MODEL NOT IMPLEMENTED YET: %s" % name)
    dims = parts[1].split("_") if len(parts) > 1 else ["512", "512"]
    d1, d2 = int(dims[0]), int(dims[1])
    dropout = "DROP" in name.split("_")
    pools = _n_pools(cfg[key])
    fh, fw = input_hw[0] // 2 ** pools, input_hw[1] // 2 ** pools
    final = [v for v in cfg[key] if v != "M"][-1]
    return VGGSlim(key, num_classes, final * fh * fw, d1, d2, dropout=dropout)


def make_alexnet(num_classes=20):
    """torchvision AlexNet (net.py:96-125) with the last layer replaced by a `num_classes` head (utils.py:68-72)."""
    import torchvision
    m = torchvision.models.alexnet(weights=None)
    m.classifier._modules["6"] = nn.Linear(4096, num_classes)
    return m


def parse_model_name(model_name, input_hw=(64, 64), num_classes=20):
    if "alexnet" in model_name:
        return make_alexnet(num_classes)
    return make_vgg(model_name, input_hw, num_classes)
