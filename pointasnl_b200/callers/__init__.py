"""PyTorch callers of the drop-in API: the two reference models that exercise the hot path
(models/pointasnl_cls.py, models/pointasnl_sem_seg_res.py), kept only as benchmark / parity drivers."""
