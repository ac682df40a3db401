* test/test_qp_io.jl:26-35: the LP above plus the objective matrix [2 2; 2 4]
NAME          trivial_qp
ROWS
 N  cost
 L  cap
COLUMNS
    x         cost      2.0   cap       1.0
    y         cost     -1.0   cap       1.0
RHS
    b         cap       3.0
BOUNDS
 UP bnd       x         1.0
 LO bnd       y         1.0
 UP bnd       y         2.0
QUADOBJ
    x         x         2.0
    x         y         2.0
    y         y         4.0
ENDATA
