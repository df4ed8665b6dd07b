"""Drop-in `simple_knn` package: the reference imports `simple_knn._C.distCUDA2`
(lightning/renderer_2dgs.py:11, lightning/point_decoder/layers/head.py:7) but does not vendor it."""
