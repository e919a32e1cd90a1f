module github.com/killingspark/sparkzstd

go 1.21
