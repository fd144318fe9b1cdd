#pragma once  /* cblas.h is included by the reference but no cblas_* call exists */
