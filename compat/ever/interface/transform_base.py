"""test-time-augmentation transforms (module/tta.py:52): transform(inputs) / inv_transform(outputs), and their product"""


class Transform(object):
    def transform(self, inputs):
        raise NotImplementedError

    def inv_transform(self, transformed_inputs):
        raise NotImplementedError


class MultiTransform(Transform):
    def __init__(self, *transforms):
        self.transforms = transforms

    def transform(self, inputs):
        return [t.transform(inputs) for t in self.transforms]

    def inv_transform(self, transformed_inputs):
        return [t.inv_transform(o) for t, o in zip(self.transforms, transformed_inputs)]
