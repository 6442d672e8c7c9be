#include <boost/numeric/ublas/matrix_sparse.hpp>
