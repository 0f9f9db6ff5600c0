# drop-in overlay package (see ../README.md)
