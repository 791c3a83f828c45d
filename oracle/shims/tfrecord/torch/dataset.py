class TFRecordDataset:
    pass
