"""Model containers with the reference's constructor signatures and state-dict keys
(reference python/models/models.py:41-62, :90-121, :124-217), so that checkpoints trained
with the reference load unchanged and ``scripts/evaluate_*.py`` can build their models from
this package.  Only what the test-time path needs is here; the training losses
(``_kld``, flows, SVI) are out of scope.

State-dict keys: ``encoder.hidden.{i}.*``, ``encoder.sample.{mu,log_var}.*``,
``decoder.hidden.{i}.*``, ``decoder.reconstruction.*``; classifier ``hidden.{i}.*``,
``output_layer.*``.  Modules are created in the same order as in the reference so that a
seeded construction draws the same initial weights.
"""
import torch
from torch import nn


def _stack(sizes):
    return nn.ModuleList(nn.Linear(a, b) for a, b in zip(sizes[:-1], sizes[1:]))


def _xavier(module):
    for m in module.modules():
        if isinstance(m, nn.Linear):
            nn.init.xavier_normal_(m.weight.data)
            if m.bias is not None:
                m.bias.data.zero_()


class GaussianSample(nn.Module):
    """Mean / log-variance heads; forward returns (sample, mu, log_var)."""

    def __init__(self, in_features, out_features):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.mu = nn.Linear(in_features, out_features)
        self.log_var = nn.Linear(in_features, out_features)

    def forward(self, x):
        mu, log_var = self.mu(x), self.log_var(x)
        noise = torch.randn(mu.size()).to(mu.device)      # drawn on the CPU generator, as upstream
        return mu.addcmul(log_var.mul(0.5).exp(), noise), mu, log_var


class Encoder(nn.Module):
    def __init__(self, dims, sample_layer=GaussianSample):
        super().__init__()
        x_dim, h_dim, z_dim = dims
        self.hidden = _stack([x_dim, *h_dim])
        self.sample = sample_layer(h_dim[-1], z_dim)

    def forward(self, x):
        for layer in self.hidden:
            x = torch.tanh(layer(x))
        return self.sample(x)


class Decoder(nn.Module):
    def __init__(self, dims):
        super().__init__()
        z_dim, h_dim, x_dim = dims
        self.hidden = _stack([z_dim, *h_dim])
        self.reconstruction = nn.Linear(h_dim[-1], x_dim)

    def forward(self, x):
        for layer in self.hidden:
            x = torch.tanh(layer(x))
        return torch.exp(self.reconstruction(x))


class Classifier(nn.Module):
    def __init__(self, dims, batch_norm=False):
        super().__init__()
        x_dim, h_dim, y_dim = dims
        if batch_norm:
            raise NotImplementedError("batch-norm classifiers are not used by the evaluate scripts")
        self.hidden = _stack([x_dim, *h_dim])
        self.output_layer = nn.Linear(h_dim[-1], y_dim)

    def forward(self, x):
        for layer in self.hidden:
            x = torch.relu(layer(x))
        return torch.sigmoid(self.output_layer(x))


class VariationalAutoencoder(nn.Module):
    """M1: dims = [x_dim, z_dim, h_dim]."""

    def __init__(self, dims):
        super().__init__()
        x_dim, z_dim, h_dim = dims
        self.z_dim = z_dim
        self.flow = None
        self.encoder = Encoder([x_dim, h_dim, z_dim])
        self.decoder = Decoder([z_dim, list(reversed(h_dim)), x_dim])
        self.kl_divergence = 0
        _xavier(self)

    def forward(self, x, y=None):
        z, mu, log_var = self.encoder(x)
        self.kl_divergence = -0.5 * torch.sum(log_var - mu.pow(2) - log_var.exp(), axis=-1)
        return self.decoder(z), mu, log_var

    def sample(self, z):
        return self.decoder(z)


class DeepGenerativeModel(VariationalAutoencoder):
    """M2: dims = [x_dim, y_dim, z_dim, h_dim]; encoder and decoder also see the label."""

    def __init__(self, dims, classifier):
        x_dim, self.y_dim, z_dim, h_dim = dims
        super().__init__([x_dim, z_dim, h_dim])
        self.encoder = Encoder([x_dim + self.y_dim, h_dim, z_dim])
        self.decoder = Decoder([z_dim + self.y_dim, list(reversed(h_dim)), x_dim])
        self.classifier = classifier
        _xavier(self)

    def forward(self, x, y):
        z, mu, log_var = self.encoder(torch.cat([x, y], dim=1))
        return self.decoder(torch.cat([z, y], dim=1)), mu, log_var

    def classify(self, x):
        return self.classifier(x)

    def sample(self, z, y):
        return self.decoder(torch.cat([z, y.float()], dim=1))
