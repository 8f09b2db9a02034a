"""Drop-in name for the reference package simple_knn (KNN/): ``from simple_knn._C import distCUDA2``
(sugar/gaussian_splatting/scene/gaussian_model.py:20)."""
