from ever.preprocess.albu import ToTensor as ToTensorV2     # noqa: F401
