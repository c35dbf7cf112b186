"""SparseConvNetTensor: features + shared Metadata + spatial size
(reference: sparseconvnet/sparseConvNetTensor.py:12-55)."""
import torch


class SparseConvNetTensor(object):
    def __init__(self, features=None, metadata=None, spatial_size=None):
        self.features = features
        self.metadata = metadata
        self.spatial_size = spatial_size

    def get_spatial_locations(self, spatial_size=None):
        """LongTensor [nActive, 4] = (x, y, z, batch index) of every row of `features`."""
        size = self.spatial_size if spatial_size is None else spatial_size
        return self.metadata.getSpatialLocations(size)

    def to(self, device):
        self.features = self.features.to(device)
        return self

    def type(self, t=None):
        if t:
            self.features = self.features.type(t)
            return self
        return self.features.type()

    def cuda(self):
        self.features = self.features.cuda()
        return self

    def cpu(self):
        self.features = self.features.cpu()
        return self

    @property
    def requires_grad(self):
        return self.features.requires_grad

    def __repr__(self):
        loc = self.get_spatial_locations() if self.metadata is not None else None
        return ("SparseConvNetTensor<<features=%r,features.shape=%r,batch_locations=%r,batch_locations.shape=%r,spatial size=%r>>"
                % (self.features, None if self.features is None else self.features.shape, loc, None if loc is None else loc.shape, self.spatial_size))
