* The LP the reference's test/test_qp_io.jl:15-24 expects from its own fixture
* (written out here from that expectation):
*   minimize 2 x - y   subject to   x + y <= 3,  0 <= x <= 1,  1 <= y <= 2
NAME          trivial_lp
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
ENDATA
